timeout 1200 python -m pytest tests/test_cpp_shim.py -x -q -m gpu > gpurun_out/t_shim.log 2>&1; echo tests rc=$?
tail -30 gpurun_out/t_shim.log | cut -c1-220
