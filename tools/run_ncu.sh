# ncu captures (run under gpurun, one GPU): launch lists and --set full captures of the hot kernels
set -x
python tools/profile_target.py c3 1 > gpurun_out/plain_c3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3.csv python tools/profile_target.py c3 1 > gpurun_out/ncu_l3.log 2>&1
python tools/profile_target.py c3 1 > gpurun_out/plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_candidates -c 2 -f -o gpurun_out/prof_c3_cand python tools/profile_target.py c3 1 > gpurun_out/ncu_c3a.log 2>&1
python tools/profile_target.py c3 1 > gpurun_out/plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"rerank|pack_operands|pack_f32" -c 6 -f -o gpurun_out/prof_c3_aux python tools/profile_target.py c3 1 > gpurun_out/ncu_c3b.log 2>&1
python tools/profile_target.py c2 1 > gpurun_out/plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv python tools/profile_target.py c2 1 > gpurun_out/ncu_l2.log 2>&1
python tools/profile_target.py c2 1 > gpurun_out/plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_candidates -c 1 -f -o gpurun_out/prof_c2_cand python tools/profile_target.py c2 1 > gpurun_out/ncu_c2.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches*
