#!/bin/bash
# round 2, call 24: device builder with presets / filter / symmetry on the GPU; default bench with the shared pinned pool
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_device_builder.py tests/test_gpu_zz_dbcsr_multiply.py -x -q 2>&1 | tail -8 | tee gpurun_out/call24_tests.log
( time timeout 900 python bench.py --no-gpu-baseline > gpurun_out/bench_r02_call24.json 2> gpurun_out/bench_r02_call24.err ) 2>&1 | tail -4
tail -c 600 gpurun_out/bench_r02_call24.err
python - <<'P'
import json
for line in open("gpurun_out/bench_r02_call24.json"):
    if line.startswith("{"):
        d = json.loads(line); e = d.get("e2e") or {}
        print("value", d["value"], "e2e", e.get("value"), e.get("ms_per_step"), e.get("stack_builder"), e.get("error"), json.dumps(e.get("variants"))[:500])
        t = d.get("tile_order") or {}; print("tile_order", t.get("value"), t.get("error"))
P
