timeout 1200 python -m pytest tests/test_gpu_fuzz.py -q -m gpu > gpurun_out/t_fuzz.log 2>&1; echo tests rc=$?
tail -40 gpurun_out/t_fuzz.log | cut -c1-220
