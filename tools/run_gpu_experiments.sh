nproc; free -g | head -2
( time python bench.py > gpurun_out/full_c3.json 2> gpurun_out/full_c3.err ) 2>&1 | grep real
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/ref_c3.json 2> gpurun_out/ref_c3.err ) 2>&1 | grep real
( time python bench.py --workload c2 > gpurun_out/full_c2.json 2> gpurun_out/full_c2.err ) 2>&1 | grep real
cat gpurun_out/ref_c3.json; tail -2 gpurun_out/ref_c3.err
python tools/bench_summary.py gpurun_out/full_c3.json gpurun_out/full_c2.json
