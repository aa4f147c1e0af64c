#!/bin/bash
# 8-GPU call after the chunk-entry kernels: C4 (FPFH, query-sharded) and C3 bench lines
mkdir -p gpurun_out/r02
O=gpurun_out/r02
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29525 bench.py --gpus 8 --steps 5 --warmup 3 --workload c4 > $O/bench_c4_8gpu_b.json 2> $O/bench_c4_8gpu_b.err; head -c 200 $O/bench_c4_8gpu_b.json; echo
NCCL_DEBUG=INFO timeout 600 $TR --master-port 29523 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_c3_8gpu_b.json 2> $O/bench_c3_8gpu_b.err; head -c 200 $O/bench_c3_8gpu_b.json; echo; grep -c "NCCL INFO" $O/bench_c3_8gpu_b.err
wc -l $O/bench_c4_8gpu_b.json $O/bench_c3_8gpu_b.json
