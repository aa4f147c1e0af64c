timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo tests rc=$?
tail -4 gpurun_out/t_all.log
( time python bench.py > gpurun_out/full_c3.json 2> gpurun_out/full_c3.err ) 2>&1 | grep real
python tools/bench_summary.py gpurun_out/full_c3.json
python -c "
import json; d=json.load(open('gpurun_out/full_c3.json')); print(json.dumps(d['roofline'], indent=1)[:1800]); print(d['e2e']); print(d['clocks'])"
