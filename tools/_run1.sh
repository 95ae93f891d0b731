timeout 300 python tests/gpu_quick.py 2>&1 | grep -v "^OK" | head -20
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu 2>&1 | tail -3
for wl in def-small atk-small def-small atk-small; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline --no-e2e --steps 300 --warmup 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl', d['ms_per_step'], d['roofline']['frac'])"
done
for v in mb7 mb8; do
  TD_B200_LIB=$PWD/build_var/$v.so timeout 300 python bench.py --workload def-small --no-cpu-baseline --no-e2e --steps 300 --warmup 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v def-small (previous commit)', d['ms_per_step'], d['roofline']['frac'])"
done
