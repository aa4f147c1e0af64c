#!/bin/bash
# 2-GPU check of bench.py's stdout handling and the e2e phase breakdown
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_c3_2gpu_c.json 2> $O/bench_c3_2gpu_c.err; wc -l $O/bench_c3_2gpu_c.json; head -c 150 $O/bench_c3_2gpu_c.json; echo; tail -3 $O/bench_c3_2gpu_c.err
