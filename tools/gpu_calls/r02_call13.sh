#!/bin/bash
# chunk-entry epilogue (EPI 4): parity of the layouts, then A/B timing on C2 / C4
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/cand_ab13.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "epilogue_layouts" > $O/pytest_gpu13.log 2>&1; tail -3 $O/pytest_gpu13.log
B200M_TC_ALT=4 timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_full_size_workloads_match_oracle_on_sampled_rows > $O/pytest_gpu13b.log 2>&1; tail -3 $O/pytest_gpu13b.log
for alt in 4 1; do
  B200M_TC_ALT=$alt timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab13.log
done
for alt in 4 2; do
  B200M_TC_ALT=$alt timeout 300 python tools/cand_time.py c4 3 2>&1 | tail -1 | tee -a $O/cand_ab13.log
done
for dbg in 1 32 256; do
  B200M_TC_ALT=4 B200M_TC_DEBUG=$dbg timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab13.log
done
B200M_TC_ALT=4 timeout 600 python tools/fullsize_parity.py c2 4096 2>&1 | tail -2
