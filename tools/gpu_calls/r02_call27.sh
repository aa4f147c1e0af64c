#!/bin/bash
# rotated sweep on C4: follower distance 16 / 20 / 24 / 28 tiles
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/cand_ab27.log
for lag in 24 16 20 28 24; do B200M_TC_SWEEP_LAG=$lag timeout 300 python tools/cand_time.py c4 3 2>&1 | tail -1 | tee -a $O/cand_ab27.log; done
