#!/bin/bash
# which SM sub-partition an epilogue warp shares with the MMA issuers: per-warp epilogue timeline (tools build), and the
# wall-clock A/B of spreading the four MMA warps over the four sub-partitions (product build)
mkdir -p gpurun_out
O=gpurun_out/r02v_spread.txt; : > $O
for sp in 0 1; do for k in 0 1 2 3; do
  echo "== spread=$sp epilogue warp $((4+k)) (sub-partition $k)" >> $O
  BIN_B200_SPREAD=$sp BIN_B200_DEBUG=$((8 + k*4096)) timeout 120 python tools/timeline.py 96 2>&1 | grep -E "^epi 2[0-3]|per tile" >> $O
done; done
timeout 900 python tools/ab_conv.py spread >> $O 2>&1
cat $O
