#!/bin/bash
# Round 2, 2-GPU call: multi-GPU inside the library -- group API (one process), rank API (torchrun), C++ shim at 2 GPUs.
mkdir -p gpurun_out/r02
O=gpurun_out/r02
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus2.txt
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_cpp_shim.py -m gpu -x -q > $O/pytest_multi2.log 2>&1; tail -4 $O/pytest_multi2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py > $O/multigpu_check_2gpu.log 2>&1; grep -E "PASS|FAIL|Error|error" $O/multigpu_check_2gpu.log | tail -30
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_c3_2gpu.json 2> $O/bench_c3_2gpu.err; tail -c 600 $O/bench_c3_2gpu.json; tail -3 $O/bench_c3_2gpu.err
