#!/bin/bash
# bench.py on a cold box, twice: does the settle phase remove the first-process penalty?
mkdir -p gpurun_out
for i in 1 2; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin.read().strip().split('\n') if l.startswith('{')][-1])
print('run $i', 'windows/s', round(j['value'],2), 'ms/step', round(j['ms_per_step'],1), 'e2e', round(j['e2e']['value'],2), j['clocks'], j["per_rank"][0]["settle"], j["per_rank"][0]["step_ms"])" >> gpurun_out/r02ag_settle.txt
done
cat gpurun_out/r02ag_settle.txt
