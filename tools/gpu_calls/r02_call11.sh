#!/bin/bash
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/cand_ab11.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_full_size_workloads_match_oracle_on_sampled_rows > $O/pytest_gpu11.log 2>&1; tail -2 $O/pytest_gpu11.log
for ne in 1 0; do
  B200M_TC_NORM_EPI=$ne timeout 300 python tools/cand_time.py c3 3 2>&1 | tail -1 | tee -a $O/cand_ab11.log
done
timeout 900 python tools/fullsize_parity.py c3 4096 2>&1 | tail -2
