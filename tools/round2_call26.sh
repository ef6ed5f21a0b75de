#!/bin/bash
# round 2, call 26: device builder with the index fetched behind the launches; e2e device leg with the pipelined left-panel upload
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_device_builder.py -x -q 2>&1 | tail -4 | tee gpurun_out/call26_tests.log
for flag in "" "--no-dev-pipelined"; do
timeout 600 python bench.py --no-gpu-baseline --no-extra --no-tiled --no-cpu --steps 5 --e2e-steps 4 $flag > gpurun_out/bench_r02_call26.json 2> gpurun_out/bench_r02_call26.err
tail -c 400 gpurun_out/bench_r02_call26.err
python - "$flag" <<'P'
import json, sys
for line in open("gpurun_out/bench_r02_call26.json"):
    if line.startswith("{"):
        d = json.loads(line); e = d.get("e2e") or {}
        print(sys.argv[1], "e2e", e.get("value"), e.get("ms_per_step"), e.get("stack_builder"), e.get("error"))
        for k, v in (e.get("variants") or {}).items(): print("  ", k, v.get("ms_per_step"), v.get("phases_last_step"), v.get("error"))
P
done
