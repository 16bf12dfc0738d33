#!/bin/bash
# epilogue ablations of the QUAD conv (tools build, timeline): which part of the ~2400-cycle epilogue is slow?
mkdir -p gpurun_out
for dbg in 8 40 72 136 264 488; do
  echo "== BIN_B200_DEBUG=$dbg (8 = none, +32 no tmem ld, +64 no bias lds, +128 no shuffles, +256 no stores, 488 = all off)" >> gpurun_out/r02i_epi_ablation.txt
  BIN_B200_QUAD=1 BIN_B200_DEBUG=$dbg timeout 200 python tools/timeline.py 96 2>&1 | grep "^epi\|per tile" | tail -n 5 >> gpurun_out/r02i_epi_ablation.txt
done
cat gpurun_out/r02i_epi_ablation.txt
