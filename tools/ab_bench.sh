#!/bin/bash
# A/B kernel timing of two builds on the same box: alternates the baseline (ipcdnnwalk_b200/_ab/libbase.so) and the
# current library, prints ms_per_step / median of each bench.py run.
for i in 1 2 3; do
  for lib in ipcdnnwalk_b200/_ab/libbase.so ipcdnnwalk_b200/libsrbtrack_b200.so; do
    IPCDNNWALK_B200_LIB=$PWD/$lib python bench.py --steps 400 --warmup 16 --no-e2e --no-cpu-baseline --no-secondary "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib', 'ms_per_step %.5f median %.5f value %.4g' % (d['ms_per_step'], d['config']['median_step_ms'], d['value']))"
  done
done
