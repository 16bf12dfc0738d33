#!/bin/bash
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests/test_gpu_conv_fuzz.py tests/test_gpu_parity.py tests/test_gpu_backward.py -q -x -p no:cacheprovider ) > gpurun_out/r02q_pytest.log 2>&1; tail -n 4 gpurun_out/r02q_pytest.log
timeout 900 python tools/ab_conv.py > gpurun_out/r02q_ab_conv.txt 2>&1; cat gpurun_out/r02q_ab_conv.txt
timeout 200 python tools/timeline.py 96 > gpurun_out/r02q_timeline_conv0_quad.txt 2>&1; grep "^epi\|per tile" gpurun_out/r02q_timeline_conv0_quad.txt | tail -n 5
