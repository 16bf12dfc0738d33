# same-box comparison of the two rdb_tail variants under the bench's CUDA-graph replay
for v in 1 0 1 0; do
  BIN_B200_TAIL_STREAMS=$v python bench.py 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('streams=$v', round(d['value'],2), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz'])"
done
