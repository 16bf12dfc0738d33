#!/bin/bash
mkdir -p gpurun_out
for b in 1 2 3 5 1; do RW_BATCH=$b timeout 300 python tools/run_window.py 6 --graph 2>&1 | tail -n 3 | sed "s/^/batch=$b /" >> gpurun_out/r02n_batch.txt; done; cat gpurun_out/r02n_batch.txt
