#!/bin/bash
# round 2, call 9: device-side stack builder -- GPU parity tests, then the e2e leg with both builders
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_device_builder.py -x -q 2>&1 | tail -25 | tee gpurun_out/call9_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu --no-gpu-baseline --e2e-steps 4 > gpurun_out/bench_r02_call9.json 2> gpurun_out/bench_r02_call9.err
tail -c 600 gpurun_out/bench_r02_call9.err
python - <<'P'
import json
for line in open("gpurun_out/bench_r02_call9.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print("value", d["value"], "e2e", json.dumps(d.get("e2e"))[:3000])
P
