#!/bin/bash
# A/B/C... kernel timing of several builds on the same box: tools/ab_multi.sh "bench args" lib1.so lib2.so ...  (3 alternating rounds)
ARGS=$1; shift
for i in 1 2 3; do
  for lib in "$@"; do
    IPCDNNWALK_B200_LIB=$PWD/$lib python bench.py --steps 400 --warmup 16 --no-e2e --no-cpu-baseline --no-secondary $ARGS 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib', 'ms_per_step %.5f median %.5f value %.4g' % (d['ms_per_step'], d['config']['median_step_ms'], d['value']))"
  done
done
