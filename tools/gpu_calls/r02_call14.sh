#!/bin/bash
# what bounds the MMA + hand-off floor of the split-N kernel: K steps per tile forced to 1..4, epilogue skipped
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/cand_ab14.log
for dbg in 3 4097 8193 12289 16385 4101 16389; do
  B200M_TC_ALT=4 B200M_TC_DEBUG=$dbg timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab14.log
done
