#!/bin/bash
# Cycle-stamp traces of the candidate kernel's accumulator hand-off on the FPFH C2 workload (run under gpurun).
# Needs a library built with the stamps compiled in:  B200M_TC_TRACE=1 python -m lidar_global_registration_b200.build
# B200M_TC_DEBUG: 1024 = stamps (general issue loop), +1 = epilogue hands the buffer back at once, +32 = TMEM drain only,
# +256 = fast path only.  Read the logs with tools/trace_report.py / tools/trace_warps.py.
for f in 1024 1025 1056 1280; do
  B200M_TC_DEBUG=$f python tools/profile_target.py c2 1 > gpurun_out/trace_c2_$f.log 2>&1
done
ls -la gpurun_out/trace_c2_*.log
