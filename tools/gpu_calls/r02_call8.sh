#!/bin/bash
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/cand_ab8.log
B200M_TC_ALT=3 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > $O/pytest_gpu8.log 2>&1; tail -2 $O/pytest_gpu8.log
for rf in 64 16 256; do
  B200M_TC_ALT=3 B200M_TC_RING_FROM=$rf timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab8.log
done
B200M_TC_ALT=3 timeout 300 python tools/cand_time.py c4 3 2>&1 | tail -1 | tee -a $O/cand_ab8.log
B200M_TC_ALT=1 timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab8.log
B200M_TC_ALT=3 timeout 600 python tools/fullsize_parity.py c2 4096 2>&1 | tail -2
