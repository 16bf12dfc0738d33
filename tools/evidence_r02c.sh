#!/bin/bash
# Round-2 evidence bundle C (run under gpurun --gpus 2): nn.DataParallel replicas on two devices, 2-rank bench (inference
# sharding + DDP training step), DDP sync check.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > gpurun_out/r02c_gpus.txt
( time timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -s -p no:cacheprovider -k "dataparallel_two" ) > gpurun_out/r02c_pytest_dp.log 2>&1
tail -n 8 gpurun_out/r02c_pytest_dp.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r02c_bench_n2.json 2> gpurun_out/r02c_bench_n2.err
tail -c 1800 gpurun_out/r02c_bench_n2.json; tail -n 5 gpurun_out/r02c_bench_n2.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/ddp_train_check.py ) > gpurun_out/r02c_ddp_check.txt 2>&1
tail -n 4 gpurun_out/r02c_ddp_check.txt
timeout 200 python tools/timeline.py 96 > gpurun_out/r02c_timeline_conv0_quad.txt 2>&1; grep "^epi\|per tile" gpurun_out/r02c_timeline_conv0_quad.txt | tail -n 6
