#!/bin/bash
# QUAD generalised to every fp16 P8 / PixelShuffle conv: numerics (fuzz + parity + backward), A/B, train step, bench
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests/test_gpu_conv_fuzz.py tests/test_gpu_parity.py tests/test_gpu_backward.py tests/test_gpu_streaming.py -q -x -p no:cacheprovider ) > gpurun_out/r02o_pytest.log 2>&1; tail -n 5 gpurun_out/r02o_pytest.log
timeout 900 python tools/ab_conv.py > gpurun_out/r02o_ab_conv.txt 2>&1; cat gpurun_out/r02o_ab_conv.txt
for q in 0 1 0 1; do BIN_B200_QUAD=$q timeout 300 python tools/bench_train.py 8 256 256 2>&1 | tail -n 1 | cut -c1-200 | sed "s/^/quad=$q /" >> gpurun_out/r02o_train.txt; done; cat gpurun_out/r02o_train.txt
