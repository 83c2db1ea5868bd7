#!/bin/bash
# usage: scripts/sweep3.sh REPS "VAR1=a VAR2=b" ...  -- REPS short bench runs per environment assignment list, medians printed
reps=$1; shift
for envs in "$@"; do
  vals=""
  for r in $(seq $reps); do
    v=$(env $envs timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); print(round(d['roofline']['frac'],4))")
    vals="$vals $v"
  done
  echo "$envs : $vals"
done
