#!/bin/bash
# final check of the round: whole GPU suite, default bench line (both CPU engines beside it), the reference arm, smoke(),
# ncu --set full of the C4 candidate kernel (forward launch) for its DRAM traffic
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu22.log 2>&1; tail -3 $O/pytest_gpu22.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > $O/bench_c3_final2.json 2> $O/bench_c3_final2.err; python tools/bench_summary.py $O/bench_c3_final2.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_c3_reference_arm2.json 2> $O/ref2.err; head -c 300 $O/bench_c3_reference_arm2.json; echo
python tools/profile_target.py c4 1 > $O/plain_c4.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_candidates -c 1 -f -o $O/prof_c4_cand python tools/profile_target.py c4 1 > $O/ncu_c4.log 2>&1
ls -la $O/prof_c4_cand.ncu-rep
