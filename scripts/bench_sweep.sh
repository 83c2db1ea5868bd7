#!/bin/bash
# usage: scripts/bench_sweep.sh VAR v1 v2 ...   -- runs the short bench once per value of an env var
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.load(sys.stdin); print('$var=$v', round(d['value'],2), 'steps/s', round(d['ms_per_step'],3), 'ms  frac', round(d['roofline']['frac'],4), d['clocks']['reasons'])"
done
