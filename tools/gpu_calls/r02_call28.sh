#!/bin/bash
# final check with the rotated sweep on by default for large train sets: whole GPU suite, smoke, default bench, C4 DRAM counters
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu28.log 2>&1; tail -3 $O/pytest_gpu28.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > $O/bench_c3_final3.json 2> $O/bench_c3_final3.err; python tools/bench_summary.py $O/bench_c3_final3.json
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:tc_candidates -c 2 --csv --log-file $O/ncu_c4_dram_final.csv python tools/profile_target.py c4 1 > $O/ncu_c4d.log 2>&1
grep -v "^==" $O/ncu_c4_dram_final.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tail -6
