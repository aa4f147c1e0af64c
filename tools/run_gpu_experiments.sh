set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "candidate_kernel_modes or tc_operands" > gpurun_out/t_tc.log 2>&1; echo tests rc=$?
tail -3 gpurun_out/t_tc.log
for w in c2 c3; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n_$w.json 2> gpurun_out/n_$w.err; echo $w rc=$?
  for d in 1 5 32; do
    B200M_TC_DEBUG=$d timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n_${w}_dbg$d.json 2> gpurun_out/n_${w}_dbg$d.err; echo $w dbg$d rc=$?
  done
done
B200M_TC_DEBUG=64 timeout 300 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n_c2_eh1.json 2> gpurun_out/n_c2_eh1.err
B200M_TC_DEBUG=128 timeout 300 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n_c3_eh2.json 2> gpurun_out/n_c3_eh2.err
python tools/bench_summary.py gpurun_out/n_*.json
