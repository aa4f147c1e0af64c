#!/bin/bash
# Round 2, GPU call 5 (1 GPU): the reworked bench.py (both arms), ncu DRAM counters of the masked reverse C3 launch.
mkdir -p gpurun_out/r02
O=gpurun_out/r02
( time python bench.py > $O/bench_c3_1gpu.json 2> $O/bench_c3_1gpu.err ) 2>&1 | grep real; tail -c 400 $O/bench_c3_1gpu.json; tail -3 $O/bench_c3_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_c3_reference_arm.json 2> $O/ref.err; tail -c 300 $O/bench_c3_reference_arm.json
python bench.py --workload c2 > $O/bench_c2_1gpu.json 2> $O/bench_c2.err; tail -c 200 $O/bench_c2_1gpu.json
python bench.py --workload c1 --no-cpu-baseline > $O/bench_c1_1gpu.json 2> $O/bench_c1.err; tail -c 200 $O/bench_c1_1gpu.json
python tools/profile_target.py c3 1 > $O/plain_c3.log 2>&1 &&
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:tc_candidates --launch-skip 1 -c 1 --csv --log-file $O/ncu_c3_launch2_dram.csv python tools/profile_target.py c3 1 > $O/ncu_c3b.log 2>&1
cat $O/ncu_c3_launch2_dram.csv | tail -5
