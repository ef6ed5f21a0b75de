#!/bin/bash
# Fifth single-GPU call: new tests (fused transpose+norms, trickle memset, 8-rank Cannon on a shared GPU), bench with memset-mode trials.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_cannon.py tests/test_gpu_smm.py tests/test_gpu_multiply.py -m gpu -q 2>&1 | tail -15
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_call5.json 2> gpurun_out/bench_r02_call5.err; tail -c 600 gpurun_out/bench_r02_call5.json; tail -5 gpurun_out/bench_r02_call5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_call5.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
print(d['config']['timed'])
c=d['clocks']; print({k:c[k] for k in c if k!='trace_every_20ms_sm_mhz_power_w'}); print(c['trace_every_20ms_sm_mhz_power_w'])
r=d['roofline']; print({k:r.get(k) for k in ('achieved','peak','frac','avg_launch_us','frac_on_timed_value','burst')})
x=d['extra_configs']
for k in x: print(k, x[k].get('value'), x[k].get('kernel_only_gflops'), x[k].get('ms_per_step'), x[k].get('zero_mode'))
PY
