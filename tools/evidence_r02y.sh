#!/bin/bash
# sustained-load A/B under the power cap: sleeping polls (BIN_B200_POLITE) judged by bench.py's own timed region + clocks
mkdir -p gpurun_out
O=gpurun_out/r02y_polite_sustained.txt; : > $O
for pol in 0 1 0 1; do
  BIN_B200_POLITE=$pol timeout 300 python bench.py --steps 30 --warmup 5 --no-extras 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('polite=$pol', 'windows/s', round(j['value'],2), 'ms/step', round(j['ms_per_step'],1), 'e2e', round(j['e2e']['value'],2), j['clocks'])" >> $O
done
cat $O
