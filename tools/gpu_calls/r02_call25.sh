#!/bin/bash
# rotated sweep: distance between followers 0 (off) / 8 / 32 / 64 tiles on C2 and C4
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/cand_ab25.log
for lag in 0 8 32 0 8; do B200M_TC_SWEEP_LAG=$lag timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab25.log; done
for lag in 0 32 64; do B200M_TC_SWEEP_LAG=$lag timeout 300 python tools/cand_time.py c4 2 2>&1 | tail -1 | tee -a $O/cand_ab25.log; done
