timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo tests rc=$?
tail -25 gpurun_out/t_all.log
