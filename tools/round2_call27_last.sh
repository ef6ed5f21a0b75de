#!/bin/bash
# last GPU call of round 2: smoke, default bench, the suites touched since the last full run
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 600 python bench.py > gpurun_out/bench_r02_last_n1.json 2> gpurun_out/bench_r02_last_n1.err ) 2>&1 | tail -4
tail -c 400 gpurun_out/bench_r02_last_n1.err
timeout 600 python -m pytest tests/test_gpu_device_builder.py tests/test_gpu_cannon.py tests/test_gpu_multiply.py tests/test_gpu_zz_dbcsr_multiply.py -x -q 2>&1 | tail -5 | tee gpurun_out/call27_tests.log
python - <<'P'
import json
for line in open("gpurun_out/bench_r02_last_n1.json"):
    if line.startswith("{"):
        d = json.loads(line); e = d.get("e2e") or {}
        print("value", d["value"], "e2e", e.get("value"), e.get("ms_per_step"), e.get("stack_builder"), e.get("error"))
        r = d["roofline"]; print(" frac", r["frac"], "burst", r["burst"]["frac"])
        t = d.get("tile_order") or {}; print(" tile_order", t.get("value"), t.get("burst_frac"), t.get("error"))
        for k, v in (d.get("extra_configs") or {}).items(): print(" ", k, v.get("value"), (v.get("selfcheck") or {}).get("ok"), v.get("error"))
        print(" gpu_baseline", (d.get("gpu_baseline") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
P
