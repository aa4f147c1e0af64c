timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo tests rc=$?
tail -12 gpurun_out/t_all.log | cut -c1-200
for w in c3 c4 c1; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n_$w.json 2> gpurun_out/n_$w.err; echo $w rc=$?; tail -2 gpurun_out/n_$w.err
done
python tools/bench_summary.py gpurun_out/n_*.json
