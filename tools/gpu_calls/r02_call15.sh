#!/bin/bash
# chunk-entry epilogues 4 (x32 loads) and 5 (x16 loads) + the compacting chunk re-rank: parity, then A/B timing on C2 / C4
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/cand_ab15.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "epilogue_layouts" > $O/pytest_gpu15.log 2>&1; tail -3 $O/pytest_gpu15.log
B200M_TC_ALT=5 timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_full_size_workloads_match_oracle_on_sampled_rows > $O/pytest_gpu15b.log 2>&1; tail -3 $O/pytest_gpu15b.log
for alt in 5 4 5 4; do
  B200M_TC_ALT=$alt timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab15.log
done
for alt in 5 4; do
  B200M_TC_ALT=$alt timeout 300 python tools/cand_time.py c4 3 2>&1 | tail -1 | tee -a $O/cand_ab15.log
done
B200M_TC_ALT=5 timeout 600 python tools/fullsize_parity.py c4 4096 2>&1 | tail -2
