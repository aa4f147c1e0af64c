# ncu captures (run under gpurun): candidate kernel on C2 and C3, re-rank kernel on C3
set -x
python tools/profile_target.py c2 1 > gpurun_out/plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_candidates -c 1 -f -o gpurun_out/prof_c2_cand python tools/profile_target.py c2 1 > gpurun_out/ncu_c2.log 2>&1
python tools/profile_target.py c3 1 > gpurun_out/plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_candidates|rerank" -c 2 -f -o gpurun_out/prof_c3 python tools/profile_target.py c3 1 > gpurun_out/ncu_c3.log 2>&1
ls -la gpurun_out/*.ncu-rep
