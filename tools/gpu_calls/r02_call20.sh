#!/bin/bash
# after the clean-up (ring epilogue removed): whole GPU suite incl. full-size parity, default bench (C3 + embedded C2 / C4 / c2_local),
# ncu launch lists of C3 and C2
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu20.log 2>&1; tail -3 $O/pytest_gpu20.log
timeout 900 python bench.py > $O/bench_c3_final.json 2> $O/bench_c3_final.err; tail -c 400 $O/bench_c3_final.json
timeout 300 python bench.py --workload c1 --steps 20 > $O/bench_c1_final.json 2> $O/bench_c1_final.err
python tools/profile_target.py c3 1 > $O/plain_c3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_c3_final.csv python tools/profile_target.py c3 1 > $O/ncu_l3.log 2>&1
python tools/profile_target.py c2 1 > $O/plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_c2_final.csv python tools/profile_target.py c2 1 > $O/ncu_l2.log 2>&1
ls -la $O/launches_c*_final.csv
