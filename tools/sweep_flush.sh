#!/bin/bash
# tuning sweep of the packed-path flush thresholds (DGCN_TC_FLUSH=early,late) on the headline shape
for f in 5,5 7,7 9,9 12,12 9,16 6,12 12,24 4,8; do
  DGCN_TC_FLUSH=$f timeout 120 python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$f', d['ms_per_step'], d['roofline']['kernel_ms'])"
done
