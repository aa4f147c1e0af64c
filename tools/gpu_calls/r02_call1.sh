#!/bin/bash
# Round 2, GPU call 1 (1 GPU): the alternating-tile epilogue (B200M_TC_ALT) -- full GPU suite, A/B timings on the FPFH
# configs, full-size parity, timing experiments, then compute-sanitizer on small cases.
mkdir -p gpurun_out/r02
O=gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
for alt in 1 0; do
  B200M_TC_ALT=$alt timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab.log
  B200M_TC_ALT=$alt timeout 300 python tools/cand_time.py c4 3 2>&1 | tail -1 | tee -a $O/cand_ab.log
done
timeout 300 python tools/cand_time.py c3 3 2>&1 | tail -1 | tee -a $O/cand_ab.log
for dbg in 1 32 256; do
  B200M_TC_DEBUG=$dbg timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab.log
  B200M_TC_ALT=0 B200M_TC_DEBUG=$dbg timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab.log
done
for sp in 1 2 3 4 6 8; do
  B200M_TC_SPLITS=$sp timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab.log
done
timeout 600 python tools/fullsize_parity.py c2 4096 2>&1 | tee $O/fullsize_parity_c2.log | tail -2
timeout 900 python tools/fullsize_parity.py c4 4096 2>&1 | tee $O/fullsize_parity_c4.log | tail -2
for tool in memcheck synccheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_target.py all > $O/sanitizer_$tool.log 2>&1
  echo "sanitizer $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_TARGET|Error|hazard" $O/sanitizer_$tool.log | head -8
done
