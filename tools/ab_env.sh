#!/bin/bash
# A/B timing of one build under two settings of an environment variable: tools/ab_env.sh VAR A B [bench args]
VAR=$1; A=$2; B=$3; shift 3
for i in 1 2 3; do
  for v in $A $B; do
    env $VAR=$v python bench.py --steps 400 --warmup 16 --no-e2e --no-cpu-baseline --no-secondary "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$VAR=$v', 'ms_per_step %.5f median %.5f value %.4g' % (d['ms_per_step'], d['config']['median_step_ms'], d['value']))"
  done
done
