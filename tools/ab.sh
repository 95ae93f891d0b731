#!/bin/bash
# A/B of library variants on one box: tools/ab.sh <workload> <steps> lib1.so lib2.so ...   (paths relative to gym_td_b200/)
wl=$1; steps=$2; shift 2
for lib in "$@"; do
  for rep in 1 2; do
    TD_B200_LIB=$PWD/gym_td_b200/$lib python bench.py --workload $wl --steps $steps --repeats 3 --no-cpu-baseline --no-e2e --no-side-workloads --replay 0 2>/dev/null \
      | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl %-28s ms/step %.4f  frac %.3f  inc %s' % ('$lib', d['ms_per_step'], d['roofline']['frac'], (d.get('incremental_obs') or {}).get('ms_per_step')))"
  done
done
