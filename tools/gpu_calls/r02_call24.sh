#!/bin/bash
# staggered rotated sweep (every CTA pair follows the most advanced one at its own distance): parity subset, C2 / C4 timing, C4 DRAM bytes
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/cand_ab24.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q --deselect tests/test_gpu_parity.py::test_full_size_workloads_match_oracle_on_sampled_rows > $O/pytest_gpu24.log 2>&1; tail -3 $O/pytest_gpu24.log
for i in 1 2; do timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab24.log; done
timeout 300 python tools/cand_time.py c4 3 2>&1 | tail -1 | tee -a $O/cand_ab24.log
timeout 600 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:tc_candidates -c 1 --csv --log-file $O/ncu_c4_dram_rot2.csv python tools/profile_target.py c4 1 > $O/ncu_c4c.log 2>&1
grep -v "^==" $O/ncu_c4_dram_rot2.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tail -3
