#!/bin/bash
# Round-2 evidence bundle D: conv MMA-warp schemes (A/B), role timelines, backward parity numbers, bench.
mkdir -p gpurun_out
timeout 900 python tools/ab_conv.py > gpurun_out/r02d_ab_conv.txt 2>&1; cat gpurun_out/r02d_ab_conv.txt
( time BIN_B200_MSPLIT=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_conv_fuzz.py tests/test_gpu_backward.py -q -p no:cacheprovider ) > gpurun_out/r02d_pytest_msplit.log 2>&1
tail -n 6 gpurun_out/r02d_pytest_msplit.log
( time timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -s -p no:cacheprovider -k "backward or weight_dist or adam" ) > gpurun_out/r02d_pytest_bwd.log 2>&1
tail -n 6 gpurun_out/r02d_pytest_bwd.log; grep -a "\[bwd\|\[weights" gpurun_out/r02d_pytest_bwd.log
timeout 300 python tools/timeline_tail.py > gpurun_out/r02d_timeline_tail.txt 2>&1; tail -n 30 gpurun_out/r02d_timeline_tail.txt
timeout 300 python tools/timeline.py 96 > gpurun_out/r02d_timeline_conv0.txt 2>&1; tail -n 25 gpurun_out/r02d_timeline_conv0.txt
BIN_B200_MSPLIT=1 timeout 300 python tools/timeline.py 96 > gpurun_out/r02d_timeline_conv0_msplit.txt 2>&1; tail -n 25 gpurun_out/r02d_timeline_conv0_msplit.txt
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
tail -c 1500 gpurun_out/r02d_bench.json
