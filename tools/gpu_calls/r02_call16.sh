#!/bin/bash
# defaults = chunk entries (5 for k <= 2, 4 beyond); chunk re-rank pass size A/B; bench lines; ncu capture of the C2 kernels
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f $O/cand_ab16.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_full_size_workloads_match_oracle_on_sampled_rows > $O/pytest_gpu16.log 2>&1; tail -3 $O/pytest_gpu16.log
for g in 1 2 4; do
  B200M_CHUNK_PASS=$g timeout 300 python tools/cand_time.py c2 10 2>&1 | tail -1 | tee -a $O/cand_ab16.log
  B200M_CHUNK_PASS=$g timeout 300 python tools/cand_time.py c4 3 2>&1 | tail -1 | tee -a $O/cand_ab16.log
done
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 > $O/bench_c2_chunk.json 2> $O/bench_c2_chunk.err; tail -c 600 $O/bench_c2_chunk.json
python tools/profile_target.py c2 1 > $O/plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_c2_chunk.csv python tools/profile_target.py c2 1 > $O/ncu_l2.log 2>&1
python tools/profile_target.py c2 1 > $O/plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_candidates|rerank" -c 2 -f -o $O/prof_c2_chunk python tools/profile_target.py c2 1 > $O/ncu_c2.log 2>&1
ls -la $O/prof_c2_chunk.ncu-rep
