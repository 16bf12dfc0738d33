#!/bin/bash
mkdir -p gpurun_out
for dbg in 8 520 8 520; do
  echo "== QUAD, BIN_B200_DEBUG=$dbg (520 = nanosleep(64) after every failed try_wait)" >> gpurun_out/r02j_polite.txt
  BIN_B200_QUAD=1 BIN_B200_DEBUG=$dbg timeout 200 python tools/timeline.py 96 2>&1 | grep "^epi\|^mma\|per tile" | tail -n 9 >> gpurun_out/r02j_polite.txt
done
cat gpurun_out/r02j_polite.txt
