#!/bin/bash
# Round-2 evidence bundle K: one mbarrier arrival per epilogue warp -- A/B of the conv schemes and pair kernels, timelines, full suite, bench
mkdir -p gpurun_out
timeout 1200 python tools/ab_conv.py > gpurun_out/r02k_ab_conv.txt 2>&1; cat gpurun_out/r02k_ab_conv.txt
timeout 600 python tools/ab_pair.py > gpurun_out/r02k_ab_pair.txt 2>&1; cat gpurun_out/r02k_ab_pair.txt
BIN_B200_QUAD=1 timeout 200 python tools/timeline.py 96 > gpurun_out/r02k_timeline_conv0_quad.txt 2>&1; grep "^epi\|per tile" gpurun_out/r02k_timeline_conv0_quad.txt | tail -n 5
timeout 200 python tools/timeline.py 96 > gpurun_out/r02k_timeline_conv0.txt 2>&1; grep "^epi\|per tile" gpurun_out/r02k_timeline_conv0.txt | tail -n 5
timeout 300 python tools/timeline_tail.py > gpurun_out/r02k_timeline_tail.txt 2>&1; tail -n 10 gpurun_out/r02k_timeline_tail.txt
( time timeout 2400 python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > gpurun_out/r02k_pytest.log 2>&1
tail -n 6 gpurun_out/r02k_pytest.log; grep -a "watchdog" gpurun_out/r02k_pytest.log
( time BIN_B200_QUAD=1 timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02k_bench_quad.json 2> gpurun_out/r02k_bench_quad.err; tail -c 600 gpurun_out/r02k_bench_quad.json
