#!/bin/bash
set -x
mkdir -p gpurun_out
P=dbcsr_b200/lib/libdbcsr_acc_b200.so
timeout 900 python -m pytest tests/test_gpu_cannon.py -m gpu -q 2>&1 | tail -5
rm -f gpurun_out/kbench_results.txt
timeout 120 ./tools/kbench $P gpurun_out 1000 0.1 20 23 0:-1:-1 2>&1 | grep -E "TFLOP|no-sync"
KBENCH_NOSYNC=1 timeout 120 ./tools/kbench $P gpurun_out 1000 0.1 20 23 0:-1:-1 2>&1 | grep -E "TFLOP|no-sync"
timeout 600 python bench.py --steps 20 --warmup 3 --no-extra --no-gpu-baseline --no-cpu --no-e2e > gpurun_out/bench_r02_call6.json 2> gpurun_out/bench_r02_call6.err; tail -5 gpurun_out/bench_r02_call6.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_call6.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
print(d['config']['timed'])
c=d['clocks']; print({k:c[k] for k in c if k!='trace_every_20ms_sm_mhz_power_w'})
r=d['roofline']; print({k:r.get(k) for k in ('achieved','peak','frac','avg_launch_us','host_enqueue_us_per_launch','frac_on_timed_value','burst')})
PY
