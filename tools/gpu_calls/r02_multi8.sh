#!/bin/bash
# Round 2, 8-GPU call: group API + rank API parity at 8 ranks, C5 (8M target rows) parity at full size, bench lines.
mkdir -p gpurun_out/r02
O=gpurun_out/r02
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_cpp_shim.py -m gpu -x -q > $O/pytest_multi8.log 2>&1; tail -3 $O/pytest_multi8.log
timeout 600 $TR --master-port 29521 tools/multigpu_check.py > $O/multigpu_check_8gpu.log 2>&1; grep -c PASS $O/multigpu_check_8gpu.log; grep -E "FAIL|Error|error" $O/multigpu_check_8gpu.log | head -5
timeout 900 $TR --master-port 29522 tools/multigpu_check.py c5 1024 > $O/multigpu_check_c5_8gpu.log 2>&1; grep -E "c5 target|Error|error" $O/multigpu_check_c5_8gpu.log | tail -3
NCCL_DEBUG=INFO timeout 600 $TR --master-port 29523 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_c3_8gpu.json 2> $O/bench_c3_8gpu.err; tail -c 300 $O/bench_c3_8gpu.json; grep -c "NCCL INFO" $O/bench_c3_8gpu.err
timeout 600 $TR --master-port 29524 bench.py --gpus 8 --steps 5 --warmup 3 --workload c5 > $O/bench_c5_8gpu.json 2> $O/bench_c5_8gpu.err; tail -c 300 $O/bench_c5_8gpu.json
timeout 600 $TR --master-port 29525 bench.py --gpus 8 --steps 5 --warmup 3 --workload c4 > $O/bench_c4_8gpu.json 2> $O/bench_c4_8gpu.err; tail -c 300 $O/bench_c4_8gpu.json
