#!/bin/bash
# 4-GPU call: Cannon parity on real NCCL ranks (2 and 4 GPUs), scaling benches N=1,2,4, reference arm at N=2.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_gpu_cannon.py -m gpu -q -k "2 or 4" 2>&1 | tail -5
timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-gpu-baseline --no-cpu --no-e2e --no-selfcheck > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; tail -2 gpurun_out/bench_r02_n1.err
for n in 2 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 20 --warmup 3 --e2e-steps 2 > gpurun_out/bench_r02_n$n.json 2> gpurun_out/bench_r02_n$n.err; tail -3 gpurun_out/bench_r02_n$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r02_ref_n2.json 2> gpurun_out/bench_r02_ref_n2.err; tail -2 gpurun_out/bench_r02_ref_n2.err
python - <<'PY'
import json
for n in (1,2,4):
    try:
        d=json.loads([l for l in open('gpurun_out/bench_r02_n%d.json'%n).read().splitlines() if l.startswith('{')][-1])
        print(n, d['value'], d['ms_per_step'], d.get('selfcheck'), (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('d2h_bytes_per_step'), d.get('exchange'), d['config'].get('parallelism'))
        if n==1: print(d['roofline']['peak'], d['roofline']['frac'], d['roofline']['burst'], d['roofline']['drain_series_after_idle_ms'])
    except Exception as ex: print(n,'FAILED',ex)
try:
    d=json.loads([l for l in open('gpurun_out/bench_r02_ref_n2.json').read().splitlines() if l.startswith('{')][-1]); print('ref', d['value'], d['cpu_baseline']['cores'], d['cpu_baseline'].get('cpu_model'))
except Exception as ex: print('ref FAILED', ex)
PY
