for cl in 1 2; do
for d in 0 1 32 256; do
    B200M_TC_CLUSTER=$cl B200M_TC_DEBUG=$d timeout 300 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n_c2_cl${cl}_dbg$d.json 2> gpurun_out/n_c2_cl${cl}_dbg$d.err; echo c2 cl$cl dbg$d rc=$?
done
done
python tools/bench_summary.py gpurun_out/n_*.json | grep -v "^  "
