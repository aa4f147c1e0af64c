#!/bin/bash
# Round 2, GPU call 2 (1 GPU): alternating-tile epilogue after the setmaxnreg fix -- GPU suite, A/B timings, full-size
# parity, train-split sweep on C3 (L2-resident train windows), ncu captures.
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/cand_ab.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
for alt in 1 0; do
  B200M_TC_ALT=$alt timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab.log
  B200M_TC_ALT=$alt timeout 300 python tools/cand_time.py c4 3 2>&1 | tail -1 | tee -a $O/cand_ab.log
done
for dbg in 1 32 256; do
  B200M_TC_DEBUG=$dbg timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab.log
done
for sp in 1 2 3 4 6 8; do
  B200M_TC_SPLITS=$sp timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab.log
done
timeout 600 python tools/fullsize_parity.py c2 4096 2>&1 | tee $O/fullsize_parity_c2.log | tail -2
timeout 900 python tools/fullsize_parity.py c4 4096 2>&1 | tee $O/fullsize_parity_c4.log | tail -2
python tools/profile_target.py c2 1 > $O/plain_c2.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_candidates -c 1 -f -o $O/prof_c2_cand python tools/profile_target.py c2 1 > $O/ncu_c2.log 2>&1
ls -la $O/*.ncu-rep
