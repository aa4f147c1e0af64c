#!/bin/bash
# 2-GPU call after the chunk-entry kernels: group API, rank API (torchrun) incl. an FPFH case, C++ shim at 2 GPUs, C4 line on 2 GPUs
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_cpp_shim.py -m gpu -x -q > $O/pytest_multi2b.log 2>&1; tail -4 $O/pytest_multi2b.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py > $O/multigpu_check_2gpu_b.log 2>&1; grep -E "PASS|FAIL|Error|error|equal|bit" $O/multigpu_check_2gpu_b.log | tail -30
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c4 --steps 3 --warmup 3 > $O/bench_c4_2gpu.json 2> $O/bench_c4_2gpu.err; tail -c 300 $O/bench_c4_2gpu.json; tail -3 $O/bench_c4_2gpu.err
