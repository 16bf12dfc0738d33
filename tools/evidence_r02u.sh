#!/bin/bash
# polite polling (wall-clock, power-capped): whole window, graph replay, with clocks
mkdir -p gpurun_out
for pol in 0 1 0 1; do
  nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader -lms 500 > /tmp/clk_$pol.txt &
  SMI=$!
  BIN_B200_POLITE=$pol timeout 300 python tools/run_window.py 40 --graph 2>&1 | tail -n 4 | sed "s/^/polite=$pol /" >> gpurun_out/r02u_polite.txt
  kill $SMI; sort -n /tmp/clk_$pol.txt | tail -n 3 | sed "s/^/polite=$pol top clocks: /" >> gpurun_out/r02u_polite.txt
done
cat gpurun_out/r02u_polite.txt
