#!/bin/bash
# Round-end measurement batch (run under gpurun, 1 GPU): tests, bench lines of every single-GPU config, the reference
# arm, full-size parity of the FPFH configs, and the ncu launch list + full capture of the FPFH candidate kernel.
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; tail -2 gpurun_out/t_all.log
bash tools/run_gpu_experiments.sh
python tools/fullsize_parity.py c2 > gpurun_out/fullsize_parity_c.log 2>&1
python tools/fullsize_parity.py c4 >> gpurun_out/fullsize_parity_c.log 2>&1
grep -c "bit-exact" gpurun_out/fullsize_parity_c.log
python tools/profile_target.py c2 1 > gpurun_out/plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv python tools/profile_target.py c2 1 > gpurun_out/ncu_l2.log 2>&1
python tools/profile_target.py c2 1 > gpurun_out/plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_candidates -c 1 -f -o gpurun_out/prof_c2_cand python tools/profile_target.py c2 1 > gpurun_out/ncu_c2.log 2>&1
ls -la gpurun_out/prof_c2_cand.ncu-rep gpurun_out/launches_c2.csv
