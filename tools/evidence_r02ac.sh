#!/bin/bash
# EPI2 + tcgen05.shift issued by the epilogue after tmem_full: parity, A/B, timeline
mkdir -p gpurun_out
O=gpurun_out/r02ac_epi2_shift.txt; : > $O
BIN_B200_EPI2=1 BIN_B200_SHIFT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "rdb_golden or rdb_shapes or window_golden" 2>&1 | tail -n 4 >> $O
timeout 900 python tools/ab_conv.py epi2 >> $O 2>&1
echo "== timeline QUAD epi2=1 shift=1" >> $O
BIN_B200_EPI2=1 BIN_B200_SHIFT=1 timeout 120 python tools/timeline.py 96 2>&1 | grep -E "^epi 2[0-3]|^mma 3[0-5]|per tile" >> $O
cut -c1-400 $O
