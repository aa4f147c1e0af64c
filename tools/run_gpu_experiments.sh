timeout 900 python -m pytest tests -x -q -m gpu -k "candidate_kernel_modes or tc_operands" > gpurun_out/t_tc.log 2>&1; echo tests rc=$?
tail -4 gpurun_out/t_tc.log
for d in 0 1 32; do
B200M_TC_DEBUG=$d timeout 600 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n_c2_dbg$d.json 2> gpurun_out/n_c2_dbg$d.err
done
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n_c4.json 2> gpurun_out/n_c4.err
timeout 600 python bench.py --workload c1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n_c1.json 2> gpurun_out/n_c1.err
timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n_c3.json 2> gpurun_out/n_c3.err
python tools/bench_summary.py gpurun_out/n_*.json
