#!/bin/bash
# usage: scripts/sweep4.sh STEPS REPS "VAR=a" ...  -- like sweep3.sh with a chosen number of timed steps (sustained-power behaviour)
steps=$1; reps=$2; shift; shift
for envs in "$@"; do
  vals=""
  for r in $(seq $reps); do
    v=$(env $envs timeout 120 python bench.py --steps $steps --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); print(round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])")
    vals="$vals | $v"
  done
  echo "$envs : $vals"
done
