#!/bin/bash
# Round-2 evidence bundle L: two MMA warps per tile stream in the fused tail (BIN_B200_TAILQ) -- numerics, A/B, timeline, window
mkdir -p gpurun_out
( time BIN_B200_TAILQ=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "rdb or window or cta_pair or backbone" ) > gpurun_out/r02l_pytest_tailq.log 2>&1; tail -n 5 gpurun_out/r02l_pytest_tailq.log
timeout 600 python tools/ab_pair.py > gpurun_out/r02l_ab_tail.txt 2>&1; cat gpurun_out/r02l_ab_tail.txt
BIN_B200_TAILQ=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 6 python tools/sanitize_pair.py tailN > gpurun_out/r02l_memcheck_tailq.txt 2>&1; grep -v "Host Frame\|^=========         in \|Saved host" gpurun_out/r02l_memcheck_tailq.txt | tail -n 3
for q in 0 1 0 1; do BIN_B200_TAILQ=$q timeout 300 python tools/run_window.py 8 --graph 2>&1 | tail -n 3 | sed "s/^/tailq=$q /" >> gpurun_out/r02l_window_tailq.txt; done; cat gpurun_out/r02l_window_tailq.txt
