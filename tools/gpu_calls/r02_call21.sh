#!/bin/bash
# average kernel rebuilt around its sequential chain: parity (the average is compared bit for bit all over the suite), C3 e2e at 1 GPU
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_full_size_workloads_match_oracle_on_sampled_rows > $O/pytest_gpu21.log 2>&1; tail -3 $O/pytest_gpu21.log
timeout 900 python bench.py --no-other-configs --no-cpu-baseline > $O/bench_c3_avg.json 2> $O/bench_c3_avg.err; python tools/bench_summary.py $O/bench_c3_avg.json
timeout 900 python bench.py --workload c2 --no-cpu-baseline --steps 10 > $O/bench_c2_avg.json 2> $O/bench_c2_avg.err; python tools/bench_summary.py $O/bench_c2_avg.json
