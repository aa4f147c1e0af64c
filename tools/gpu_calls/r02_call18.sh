#!/bin/bash
# mbarrier waits without the printf on the time-out path: issuer loop 126 -> 81 instructions per tile
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/cand_ab18.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_full_size_workloads_match_oracle_on_sampled_rows > $O/pytest_gpu18.log 2>&1; tail -3 $O/pytest_gpu18.log
for dbg in 0 0 1 3 32 256; do
  B200M_TC_DEBUG=$dbg timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab18.log
done
B200M_TC_ALT=4 timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab18.log
timeout 300 python tools/cand_time.py c4 3 2>&1 | tail -1 | tee -a $O/cand_ab18.log
timeout 300 python tools/cand_time.py c3 3 2>&1 | tail -1 | tee -a $O/cand_ab18.log
