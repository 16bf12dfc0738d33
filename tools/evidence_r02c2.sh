#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r02c2_bench_n2.json 2> gpurun_out/r02c2_bench_n2.err
grep -a "^{" gpurun_out/r02c2_bench_n2.json | head -c 700; tail -n 4 gpurun_out/r02c2_bench_n2.err
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras ) > gpurun_out/r02c2_bench_n1.json 2> gpurun_out/r02c2_bench_n1.err
grep -a "^{" gpurun_out/r02c2_bench_n1.json | head -c 500
