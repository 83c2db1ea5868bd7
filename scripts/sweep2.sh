#!/bin/bash
# usage: scripts/sweep2.sh "VAR1=a VAR2=b" "VAR1=c" ...  -- one short bench run per environment assignment list
for envs in "$@"; do
  env $envs python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.load(sys.stdin); print('$envs', round(d['value'],2), 'steps/s', round(d['ms_per_step'],3), 'ms  frac', round(d['roofline']['frac'],4), d['clocks']['reasons'])"
done
