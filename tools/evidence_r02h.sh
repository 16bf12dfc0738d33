#!/bin/bash
# Round-2 evidence bundle H: four-MMA-warp (QUAD) conv scheme -- numerics, A/B, timeline
mkdir -p gpurun_out
( time BIN_B200_QUAD=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_conv_fuzz.py -q -x -p no:cacheprovider ) > gpurun_out/r02h_pytest_quad.log 2>&1; tail -n 5 gpurun_out/r02h_pytest_quad.log
timeout 1200 python tools/ab_conv.py > gpurun_out/r02h_ab_conv.txt 2>&1; cat gpurun_out/r02h_ab_conv.txt
BIN_B200_QUAD=1 timeout 300 python tools/timeline.py 96 > gpurun_out/r02h_timeline_conv0_quad.txt 2>&1; grep "^mma" gpurun_out/r02h_timeline_conv0_quad.txt | sed -n 12,20p; tail -n 1 gpurun_out/r02h_timeline_conv0_quad.txt
BIN_B200_QUAD=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 6 python tools/sanitize_pair.py convN > gpurun_out/r02h_memcheck_quad.txt 2>&1; grep -v "Host Frame\|^=========         in \|Saved host" gpurun_out/r02h_memcheck_quad.txt | tail -n 4
