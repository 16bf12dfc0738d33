#!/bin/bash
# x-stack sum by tcgen05.shift: bit-identity across the switches, A/B, timeline
mkdir -p gpurun_out
O=gpurun_out/r02x_shift.txt; : > $O
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_identical or rdb or conv" 2>&1 | tail -n 5 >> $O
timeout 900 python tools/ab_conv.py shift >> $O 2>&1
python -m bin_b200.build --tools > /dev/null 2>&1
for sh in 0 1; do
  echo "== timeline QUAD shift=$sh" >> $O
  BIN_B200_SHIFT=$sh timeout 120 python tools/timeline.py 96 2>&1 | grep -E "^epi 2[0-3]|^mma 3[0-5]|per tile" >> $O
done
cat $O
