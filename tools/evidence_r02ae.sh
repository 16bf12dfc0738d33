#!/bin/bash
# bench.py on a cold box, twice: settle phase + per-step times of the timed region
mkdir -p gpurun_out
for i in 1 2; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r02ai_bench_$i.json 2>/dev/null
done
python tools/show_bench.py gpurun_out/r02ai_bench_1.json gpurun_out/r02ai_bench_2.json | tee gpurun_out/r02ai_settle.txt
