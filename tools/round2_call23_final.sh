#!/bin/bash
# final 1-GPU call of round 2: full GPU suite, smoke, default bench + reference arm, ncu launch list of the bench command
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6 ) 2>&1 | tee gpurun_out/call23_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/bench_r02_final_n1.json 2> gpurun_out/bench_r02_final_n1.err ) 2>&1 | tail -4
tail -c 500 gpurun_out/bench_r02_final_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_r02_final_n1_ref.json 2> gpurun_out/bench_r02_final_n1_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-extra --no-e2e --no-cpu --no-gpu-baseline --no-tiled --no-selfcheck --no-clock-sampler > gpurun_out/ncu_launch_list.log 2>&1
for st in 1 2 8; do
  timeout 300 python bench.py --config cfg3 --cfg3-streams $st --steps 10 --warmup 3 --no-e2e --no-cpu --no-selfcheck --no-clock-sampler 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('cfg3 streams $st value', d['value'], 'kernel_only', d['roofline']['kernel_only_gflops'])"
done
python - <<'P'
import csv, json, collections
for f in ("gpurun_out/bench_r02_final_n1.json", "gpurun_out/bench_r02_final_n1_ref.json"):
    for line in open(f):
        if line.startswith("{"):
            d = json.loads(line)
            e = d.get("e2e") or {}
            print(f, "value", d["value"], "e2e", e.get("value"), e.get("ms_per_step"), e.get("stack_builder"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
            if "roofline" in d:
                r = d["roofline"]; print(" frac", r["frac"], "burst", r["burst"]["frac"])
                t = d.get("tile_order") or {}; print(" tile_order", t.get("value"), t.get("kernel_only_gflops"), t.get("burst_frac"), t.get("error"))
                for k, v in (d.get("extra_configs") or {}).items(): print(" ", k, v.get("value"), (v.get("selfcheck") or {}).get("ok"), v.get("error"))
rows = [r for r in csv.reader(open("gpurun_out/r02_ncu_launches_bench.csv", errors="replace")) if len(r) > 10]
h = rows[0]; iv, iname = h.index("Metric Value"), h.index("Kernel Name")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    k = r[iname].split("(")[0][:60]; tot[k] += float(r[iv].replace(",", "")); cnt[k] += 1
s = sum(tot.values())
for k, v in tot.most_common(6): print("%-62s %5d launches %8.3f ms %5.1f %%" % (k, cnt[k], v * 1e-6, 100 * v / s))
P
