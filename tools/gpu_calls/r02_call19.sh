#!/bin/bash
# quarter columns + chunk entries (EPI 7): parity of the layouts, A/B against EPI 5 / 4
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/cand_ab19.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "epilogue_layouts" > $O/pytest_gpu19.log 2>&1; tail -3 $O/pytest_gpu19.log
for alt in 7 5 7 5; do
  B200M_TC_ALT=$alt timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab19.log
done
for dbg in 1 32 256; do
  B200M_TC_ALT=7 B200M_TC_DEBUG=$dbg timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab19.log
done
for alt in 7 4; do
  B200M_TC_ALT=$alt timeout 300 python tools/cand_time.py c4 3 2>&1 | tail -1 | tee -a $O/cand_ab19.log
done
