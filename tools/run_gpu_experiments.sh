( time python bench.py > gpurun_out/full_c3.json 2> gpurun_out/full_c3.err ) 2>&1 | grep real
python bench.py --workload c2 > gpurun_out/full_c2.json 2> gpurun_out/full_c2.err
python bench.py --workload c4 --steps 3 --no-cpu-baseline > gpurun_out/full_c4.json 2> gpurun_out/full_c4.err
python bench.py --workload c1 --no-cpu-baseline > gpurun_out/full_c1.json 2> gpurun_out/full_c1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/ref_c3.json 2> gpurun_out/ref_c3.err
python tools/bench_summary.py gpurun_out/full_c3.json gpurun_out/full_c2.json gpurun_out/full_c4.json gpurun_out/full_c1.json
python -c "
import json; d=json.load(open('gpurun_out/ref_c3.json')); print('reference arm', d['value'], d['cpu_baseline']['cores'])"
