#!/bin/bash
# 2-GPU call: Cannon test with the distributed-input case over NCCL, bench --gpus 2 with the region timing
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cannon.py -m gpu -q -k "2" 2>&1 | tail -15 | tee gpurun_out/call20_tests.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 20 --warmup 3 --e2e-steps 2 > gpurun_out/bench_r02_region_n2.json 2> gpurun_out/bench_r02_region_n2.err; tail -3 gpurun_out/bench_r02_region_n2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r02_region_n2.json').read().splitlines() if l.startswith('{')][-1])
print(2, d['value'], d['ms_per_step'], d['config'].get('isolated_ms_per_step'), (d.get('selfcheck') or {}).get('ok'), (d.get('e2e') or {}).get('value'))
PY
