#!/bin/bash
# round 2, call 17: what the driver runs at round end on one GPU -- full GPU suite, smoke, default bench, reference arm
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/call17_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
( time timeout 900 python bench.py > gpurun_out/bench_r02_call17.json 2> gpurun_out/bench_r02_call17.err ) 2>&1 | tail -4
tail -c 800 gpurun_out/bench_r02_call17.err
( time timeout 600 python bench.py --impl reference > gpurun_out/bench_r02_call17_ref.json 2> gpurun_out/bench_r02_call17_ref.err ) 2>&1 | tail -4
python - <<'P'
import json
for f in ("gpurun_out/bench_r02_call17.json", "gpurun_out/bench_r02_call17_ref.json"):
    for line in open(f):
        if line.startswith("{"):
            d = json.loads(line)
            e = d.get("e2e") or {}
            print(f, "value", d["value"], "e2e", e.get("value"), e.get("ms_per_step"), e.get("stack_builder"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
            if "roofline" in d:
                r = d["roofline"]; print(" frac", r["frac"], "burst", r["burst"]["frac"], "dgemm", r.get("cublas_dgemm_8192_gflops"), r.get("cublas_dgemm_8192_sustained_gflops"))
                t = d.get("tile_order") or {}; print(" tile_order", t.get("value"), t.get("kernel_only_gflops"), t.get("burst_frac"), t.get("error"))
                x = d.get("extra_configs") or {}
                for k, v in x.items(): print(" ", k, v.get("value"), v.get("kernel_only_gflops"), (v.get("selfcheck") or {}).get("ok"), v.get("error"))
                print(" gpu_baseline", (d.get("gpu_baseline") or {}).get("value"), "variants", json.dumps(e.get("variants"))[:600])
P
