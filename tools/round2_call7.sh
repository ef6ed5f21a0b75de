#!/bin/bash
set -x
mkdir -p gpurun_out
for flag in "" "--no-clock-sampler"; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-extra --no-gpu-baseline --no-cpu --no-e2e --no-selfcheck $flag > gpurun_out/bench_r02_call7.json 2> gpurun_out/bench_r02_call7.err; tail -3 gpurun_out/bench_r02_call7.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_call7.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')})
r=d['roofline']; print({k:r.get(k) for k in ('achieved','avg_launch_us','host_enqueue_us_per_launch','drain_series_after_idle_ms')}, r['burst']['kernel_only_gflops'])
print(d['config']['timed'][-260:])
PY
done
