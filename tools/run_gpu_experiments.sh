timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo tests rc=$?
tail -15 gpurun_out/t_all.log
for w in c3; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n_$w.json 2> gpurun_out/n_$w.err; echo $w rc=$?; tail -2 gpurun_out/n_$w.err
  B200M_TC_AUG=1 timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n_${w}_aug1.json 2> gpurun_out/n_${w}_aug1.err; echo $w rc=$?; tail -2 gpurun_out/n_${w}_aug1.err
done
timeout 600 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n_c2.json 2> gpurun_out/n_c2.err
python tools/bench_summary.py gpurun_out/n_*.json
