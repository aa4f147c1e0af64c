#!/bin/bash
# Round 2, GPU call 3 (1 GPU): quarter-column epilogue (B200M_TC_ALT=2) vs alternating tiles (1) vs eight warps (0); the
# wide-seam matcher tests; timing experiments incl. "no operand traffic" (flag 4) to see what L2 -> SM bandwidth costs.
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/cand_ab3.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu3.log 2>&1; tail -3 $O/pytest_gpu3.log
for alt in 2 1 0; do
  B200M_TC_ALT=$alt timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab3.log
  B200M_TC_ALT=$alt timeout 300 python tools/cand_time.py c4 3 2>&1 | tail -1 | tee -a $O/cand_ab3.log
done
for dbg in 1 5 32 256 260; do
  B200M_TC_DEBUG=$dbg timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab3.log
done
B200M_TC_ALT=1 B200M_TC_DEBUG=5 timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab3.log
B200M_TC_ALT=0 B200M_TC_DEBUG=5 timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab3.log
timeout 300 python tools/cand_time.py c1 20 2>&1 | tail -1 | tee -a $O/cand_ab3.log
timeout 600 python tools/fullsize_parity.py c2 4096 2>&1 | tee $O/fullsize_parity_c2.log | tail -2
timeout 900 python tools/fullsize_parity.py c4 4096 2>&1 | tee $O/fullsize_parity_c4.log | tail -2
