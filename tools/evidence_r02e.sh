#!/bin/bash
# Round-2 evidence bundle E: single-lane barrier polling -- A/B of the MMA-warp schemes and the pair kernels, full GPU suite, timelines, bench.
mkdir -p gpurun_out
timeout 1200 python tools/ab_conv.py > gpurun_out/r02e_ab_conv.txt 2>&1; cat gpurun_out/r02e_ab_conv.txt
timeout 600 python tools/ab_pair.py > gpurun_out/r02e_ab_pair.txt 2>&1; cat gpurun_out/r02e_ab_pair.txt
( time timeout 2400 python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > gpurun_out/r02e_pytest.log 2>&1
tail -n 8 gpurun_out/r02e_pytest.log; grep -a "watchdog\|\[bwd\|\[fullsize\|\[weights" gpurun_out/r02e_pytest.log
timeout 300 python tools/timeline_tail.py > gpurun_out/r02e_timeline_tail.txt 2>&1; tail -n 24 gpurun_out/r02e_timeline_tail.txt
timeout 300 python tools/timeline.py 96 > gpurun_out/r02e_timeline_conv0.txt 2>&1; grep "^mma" gpurun_out/r02e_timeline_conv0.txt | sed -n 12,20p; tail -n 1 gpurun_out/r02e_timeline_conv0.txt
BIN_B200_MSPLIT=1 timeout 300 python tools/timeline.py 96 > gpurun_out/r02e_timeline_conv0_msplit.txt 2>&1; grep "^mma" gpurun_out/r02e_timeline_conv0_msplit.txt | sed -n 12,20p; tail -n 1 gpurun_out/r02e_timeline_conv0_msplit.txt
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
tail -c 1200 gpurun_out/r02e_bench.json
