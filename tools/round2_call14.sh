#!/bin/bash
# round 2, call 14: BF16 planned kernel with two issuer warps; FP64 kernel with 4-entry alignment look-ahead (reference + tile order)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16_tiled.py tests/test_gpu_smm.py -x -q 2>&1 | tail -8 | tee gpurun_out/call14_tests.log
timeout 400 python bench.py --config cfg4 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_cfg4_v4.json 2> gpurun_out/bench_cfg4_v4.err
tail -c 300 gpurun_out/bench_cfg4_v4.err
python - <<'P'
import json
for line in open("gpurun_out/bench_cfg4_v4.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print("cfg4 value", d["value"], "ms", d["ms_per_step"], "selfcheck", d.get("selfcheck"))
P
timeout 500 python bench.py --steps 20 --warmup 3 --no-extra --no-e2e --no-cpu --no-gpu-baseline > gpurun_out/bench_r02_call14.json 2> gpurun_out/bench_r02_call14.err
tail -c 600 gpurun_out/bench_r02_call14.err
python - <<'P'
import json
for line in open("gpurun_out/bench_r02_call14.json"):
    if line.startswith("{"):
        d = json.loads(line); r = d["roofline"]
        print("value", d["value"], "kernel_only", r["kernel_only_gflops"], "burst", r["burst"]["kernel_only_gflops"], "series", r["drain_series_after_idle_ms"])
        t = d.get("tile_order") or {}
        print("tile_order", {k: t.get(k) for k in ("value", "kernel_only_gflops", "burst_kernel_only_gflops", "drain_series_after_idle_ms", "zero_mode", "error")})
P
