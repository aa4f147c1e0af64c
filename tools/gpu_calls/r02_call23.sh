#!/bin/bash
# rotated sweep start of the chunk-entry kernels (CTA pairs start where the running ones are): parity, C2 / C4 timing, DRAM bytes of C4
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/cand_ab23.log
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu23.log 2>&1; tail -3 $O/pytest_gpu23.log
for i in 1 2; do timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab23.log; done
timeout 300 python tools/cand_time.py c4 3 2>&1 | tail -1 | tee -a $O/cand_ab23.log
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:tc_candidates -c 2 --csv --log-file $O/ncu_c4_dram_rot.csv python tools/profile_target.py c4 1 > $O/ncu_c4b.log 2>&1
grep -v "^==" $O/ncu_c4_dram_rot.csv | cut -d, -f5,10-15 | tail -7
timeout 900 python bench.py --workload c4 --steps 3 --no-cpu-baseline > $O/bench_c4_rot.json 2> $O/bench_c4_rot.err; python tools/bench_summary.py $O/bench_c4_rot.json
