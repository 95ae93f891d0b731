#!/bin/bash
# One GPU-box call that refreshes the round's measurements (run under gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash tools/round_measure.sh'
# GPU tests, the default bench line (with e2e, cpu_baseline, reference arm), every other workload, and the
# ncu launch list of the bench command (after the same command exited 0 without ncu).  Outputs in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python bench.py > $O/bench_r01.json 2> $O/bench_r01.err; tail -c 600 $O/bench_r01.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r01_reference.json 2> $O/bench_r01_reference.err
for wl in atk-small def-middle-multi def-middle-multi-sparse 2p-large def-middle def-large; do
  python bench.py --workload $wl --no-cpu-baseline --steps 100 > $O/bench_$wl.json 2> $O/bench_$wl.err
done
CMD="python bench.py --steps 5 --warmup 3 --preroll 1300 --no-cpu-baseline --no-e2e"
$CMD > $O/launch_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 12 --csv --log-file $O/launches.csv $CMD > $O/launch_ncu.log 2>&1
echo "launch list rc=$?"
python -c "
import json,glob
for f in sorted(glob.glob('$O/bench_*.json')):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f, 'unreadable'); continue
    e=d.get('e2e') or {}
    print('%-36s %-9s value %.4g %s  ms/step %.4f  roofline %s  e2e %s' % (f.split('/')[-1], d.get('impl','b200'), d['value'], d['unit'], d['ms_per_step'], (d.get('roofline') or {}).get('frac'), e.get('value')))
" | tee $O/bench_summary.txt
