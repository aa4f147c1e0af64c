#!/bin/bash
# 2-GPU parity with the rotated sweep forced on (small cases would not use it by default)
mkdir -p gpurun_out/r02
O=gpurun_out/r02
B200M_TC_SWEEP_LAG=5 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py > $O/multigpu_check_2gpu_rot.log 2>&1; grep -c PASS $O/multigpu_check_2gpu_rot.log; grep -E "FAIL|Error|error" $O/multigpu_check_2gpu_rot.log | head -5
B200M_TC_SWEEP_LAG=5 timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
