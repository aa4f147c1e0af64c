#!/bin/bash
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/cand_ab6.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu6.log 2>&1; tail -2 $O/pytest_gpu6.log
for alt in 1 2 0; do
  B200M_TC_ALT=$alt timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab6.log
  B200M_TC_ALT=$alt timeout 300 python tools/cand_time.py c4 3 2>&1 | tail -1 | tee -a $O/cand_ab6.log
done
timeout 600 python tools/fullsize_parity.py c2 4096 2>&1 | tail -2
