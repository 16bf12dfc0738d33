#!/bin/bash
# scaling sanity on one 8-GPU node: N = 8 then N = 4 (bench.py under torchrun, as the driver launches it)
mkdir -p gpurun_out
for n in 8 4; do
  ( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 ) > gpurun_out/r02s_bench_n$n.json 2> gpurun_out/r02s_bench_n$n.err
  grep -a "^{" gpurun_out/r02s_bench_n$n.json | head -c 300; echo; tail -n 3 gpurun_out/r02s_bench_n$n.err
done
