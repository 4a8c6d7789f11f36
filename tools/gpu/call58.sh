#!/bin/bash
# round 2, call 58: uploads in chunks on two copy streams vs one copy per tensor on one stream: pipeline tests, e2e A/B
timeout 1200 python -m pytest tests/test_gpu_model.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -2
for rep in 1 2 3; do for n in 1 2; do
  RP_COPY_STREAMS=$n timeout 900 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('copy_streams=$n value', round(d['value']), 'e2e', round(e['value']), 'e2e ms', round(e['ms_per_step'],3), 'bf16 rows', round(e['bf16_feature_rows']['value']), d['clocks']['sm_mhz'])"
done; done
