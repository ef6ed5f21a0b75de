#!/bin/bash
# round 2, call 25: operator-level tests with the device-side builder; clean ncu launch list of the bench command
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zz_dbcsr_multiply.py -x -q -k "device_builder" 2>&1 | tail -6 | tee gpurun_out/call25_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-extra --no-e2e --no-cpu --no-gpu-baseline --no-tiled --no-selfcheck --no-clock-sampler --no-peak-probes > gpurun_out/ncu_launch_list.log 2>&1
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_ncu_launches_bench.csv", errors="replace")) if len(r) > 10]
h = rows[0]; iv, iname = h.index("Metric Value"), h.index("Kernel Name")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    k = r[iname].split("(")[0][:70]; tot[k] += float(r[iv].replace(",", "")); cnt[k] += 1
s = sum(tot.values())
for k, v in tot.most_common(6): print("%-72s %5d launches %8.3f ms %5.1f %%" % (k, cnt[k], v * 1e-6, 100 * v / s))
P
