#!/bin/bash
# 8-GPU call: Cannon parity incl. distributed input on real NCCL ranks (2, 4, 8), scaling benches N=8,4,2 with the contract's region timing (C replay step)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cannon.py -m gpu -q 2>&1 | tail -12 | tee gpurun_out/call22_tests.log
for n in 8 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 20 --warmup 3 --e2e-steps 2 > gpurun_out/bench_r02_region_n$n.json 2> gpurun_out/bench_r02_region_n$n.err; tail -2 gpurun_out/bench_r02_region_n$n.err | cut -c1-300
done
python - <<'PY'
import json
for n in (2,4,8):
    try:
        d=json.loads([l for l in open('gpurun_out/bench_r02_region_n%d.json'%n).read().splitlines() if l.startswith('{')][-1])
        print(n, d['value'], d['ms_per_step'], 'isolated', d['config'].get('isolated_ms_per_step'), (d.get('selfcheck') or {}).get('ok'), 'e2e', (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('ms_per_step'), d.get('exchange'), d.get('cuda_graph'))
    except Exception as ex: print(n,'FAILED',ex)
PY
