for wl in def-small atk-small; do
  CMD="python bench.py --workload $wl --no-cpu-baseline --no-e2e --steps 3 --warmup 3 --preroll 1300"
  $CMD > gpurun_out/plain_$wl.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:td_step_kernel -s 1303 -c 2 -f -o gpurun_out/r01c_step_$wl $CMD > gpurun_out/ncu_$wl.log 2>&1
  echo "rc=$?"
done
