#!/bin/bash
# 8-GPU call: Cannon parity on real NCCL ranks (2, 4, 8 GPUs), scaling benches N=2,4,8 (N=1: bench_r02_full_n1.json), reference arm at N=8
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_gpu_cannon.py -m gpu -q 2>&1 | tail -5 | tee gpurun_out/call19_tests.log
for n in 8 4 2; do
  DBCSR_B200_CANNON_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 20 --warmup 3 --e2e-steps 2 > gpurun_out/bench_r02_final_n$n.json 2> gpurun_out/bench_r02_final_n$n.err; tail -3 gpurun_out/bench_r02_final_n$n.err | cut -c1-300
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/bench_r02_final_ref_n8.json 2> gpurun_out/bench_r02_final_ref_n8.err; tail -2 gpurun_out/bench_r02_final_ref_n8.err
python - <<'PY'
import json
for n in (2,4,8):
    try:
        d=json.loads([l for l in open('gpurun_out/bench_r02_final_n%d.json'%n).read().splitlines() if l.startswith('{')][-1])
        print(n, d['value'], d['ms_per_step'], d.get('selfcheck'), (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('ms_per_step'), d.get('exchange'), d.get('cuda_graph'), d.get('graph_capture_error'))
    except Exception as ex: print(n,'FAILED',ex)
try:
    d=json.loads([l for l in open('gpurun_out/bench_r02_final_ref_n8.json').read().splitlines() if l.startswith('{')][-1]); print('ref', d['value'], d['cpu_baseline']['cores'], d['cpu_baseline'].get('cpu_model'))
except Exception as ex: print('ref FAILED', ex)
PY
