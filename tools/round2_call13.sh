#!/bin/bash
# round 2, call 13: BF16 planned kernel vs pipeline depth; launch-policy sweep on tile-ordered FP64 stacks
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
for st in 5 4 3; do
  DBCSR_B200_BF16_STAGES=$st timeout 400 python bench.py --config cfg4 --steps 4 --warmup 3 --no-e2e --no-cpu --no-selfcheck \
    > gpurun_out/bench_cfg4_st$st.json 2> gpurun_out/bench_cfg4_st$st.err
  tail -c 300 gpurun_out/bench_cfg4_st$st.err
  python - "$st" <<'P'
import json, sys
for line in open("gpurun_out/bench_cfg4_st%s.json" % sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        print("stages", sys.argv[1], "value", d["value"], "ms", d["ms_per_step"])
P
done
for tile in 64 128; do
timeout 500 python bench.py --steps 6 --warmup 3 --no-extra --no-e2e --no-cpu --no-gpu-baseline --no-selfcheck --tiled-sweep --dev-tile $tile > gpurun_out/bench_r02_call13_t$tile.json 2> gpurun_out/bench_r02_call13_t$tile.err
tail -c 600 gpurun_out/bench_r02_call13_t$tile.err
python - "$tile" <<'P'
import json, sys
for line in open("gpurun_out/bench_r02_call13_t%s.json" % sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        t = d.get("tile_order") or {}
        print("tile", sys.argv[1], "value", t.get("value"), "kernel_only", t.get("kernel_only_gflops"), "sweep", json.dumps(t.get("policy_sweep_ms_burst_sustained")))
P
done
