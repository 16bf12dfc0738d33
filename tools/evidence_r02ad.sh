#!/bin/bash
# 2-GPU check of the final tree: DataParallel replica test + the N=2 bench exactly as the driver launches it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k "dataparallel or data_parallel or replica" 2>&1 | tail -n 3 > gpurun_out/r02ad_dp.txt
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r02ad_bench_n2.json 2> gpurun_out/r02ad_bench_n2.err
cat gpurun_out/r02ad_dp.txt; tail -c 1500 gpurun_out/r02ad_bench_n2.json | cut -c1-1500; tail -n 4 gpurun_out/r02ad_bench_n2.err
