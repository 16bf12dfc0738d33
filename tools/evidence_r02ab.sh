#!/bin/bash
# two alternating epilogue sets in the x-stacked conv (BIN_B200_EPI2): parity, A/B, timeline
mkdir -p gpurun_out
O=gpurun_out/r02ab_epi2.txt; : > $O
BIN_B200_EPI2=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rdb or golden" 2>&1 | tail -n 3 >> $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_identical" 2>&1 | tail -n 3 >> $O
timeout 900 python tools/ab_conv.py epi2 >> $O 2>&1
for e in 0 1; do
  echo "== timeline QUAD epi2=$e" >> $O
  BIN_B200_EPI2=$e timeout 120 python tools/timeline.py 96 2>&1 | grep -E "^epi 2[0-3]|^mma 3[0-5]|per tile" >> $O
done
cut -c1-400 $O
