#!/bin/bash
# One GPU-box call that refreshes the round's measurements (run under gpurun from the repo root):
#   gpurun --timeout 2400 -- 'bash tools/round_measure.sh r02'
# GPU tests, the default bench line (headline + e2e + replay + side workloads + cpu_baseline), the reference arm,
# the ncu launch list of the bench command and one `ncu --set full` capture per benchmark workload (each after the
# same command exited 0 without ncu).  Outputs in gpurun_out/.
set -u
R=${1:-r02}
O=gpurun_out
mkdir -p $O
if [ "${SKIP_TESTS:-0}" != "1" ]; then python -m pytest tests -x -q -m gpu > $O/${R}_pytest_gpu.log 2>&1; tail -2 $O/${R}_pytest_gpu.log; fi
python bench.py > $O/${R}_bench.json 2> $O/${R}_bench.err; tail -c 400 $O/${R}_bench.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_reference.json 2> $O/${R}_bench_reference.err
B="--steps 5 --warmup 3 --repeats 1 --preroll 1300 --no-cpu-baseline --no-e2e --no-side-workloads --replay 0"
CMD="python bench.py $B"
$CMD > $O/${R}_launch_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 12 --csv --log-file $O/${R}_launches.csv $CMD > $O/${R}_launch_ncu.log 2>&1
echo "launch list rc=$?"
for wl in def-small atk-small def-middle-multi 2p-large; do
  CMD="python bench.py --workload $wl --steps 3 --warmup 3 --repeats 1 --preroll 1300 --no-cpu-baseline --no-e2e --no-side-workloads --replay 0"
  $CMD > $O/${R}_plain_$wl.log 2>&1 && \
  ncu --set full --clock-control none -k regex:td_step_kernel -s 1305 -c 1 -f -o $O/${R}_step_$wl $CMD > $O/${R}_ncu_$wl.log 2>&1
  echo "ncu $wl rc=$?"
  # gpurun brings back at most 64 MiB: keep the text summaries of every capture, the reports of the two 10x10 workloads
  ncu -i $O/${R}_step_$wl.ncu-rep --page raw --csv > $O/${R}_step_${wl}_raw.csv 2>/dev/null
  ncu -i $O/${R}_step_$wl.ncu-rep --page details > $O/${R}_step_${wl}_details.txt 2>/dev/null
  python tools/ncu_sass_top.py $O/${R}_step_$wl.ncu-rep > $O/${R}_step_${wl}_sass_top.txt 2>/dev/null
  case $wl in def-small|atk-small) ;; *) rm -f $O/${R}_step_$wl.ncu-rep ;; esac
done
python tools/e2e_sweep.py def-small > $O/${R}_e2e_sweep.txt 2>&1; head -4 $O/${R}_e2e_sweep.txt
