#!/bin/bash
# rotated sweep on C4: follower distance 0 (off) against 32 / 24 / 48 tiles, alternating on one box
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/cand_ab26.log
for lag in 0 32 24 0 32 48 0 32; do B200M_TC_SWEEP_LAG=$lag timeout 300 python tools/cand_time.py c4 3 2>&1 | tail -1 | tee -a $O/cand_ab26.log; done
