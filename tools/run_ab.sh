#!/bin/bash
for f in 256 131328 262400; do
  B200M_TC_DEBUG=$f python bench.py --workload c2 --no-cpu-baseline --steps 5 > gpurun_out/ab_c2_f$f.json 2> gpurun_out/ab_c2_f$f.err
done
grep -o '"ms_candidates": [0-9.]*' gpurun_out/ab_c2_f256.json gpurun_out/ab_c2_f131328.json gpurun_out/ab_c2_f262400.json
