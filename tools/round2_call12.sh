#!/bin/bash
# round 2, call 12: planned BF16 kernel with smem-staged run words; tile-order leg of the FP64 bench; device-builder tests
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16_tiled.py tests/test_gpu_device_builder.py -x -q 2>&1 | tail -15 | tee gpurun_out/call12_tests.log
for mode in "1 0" "1 1"; do
  set -- $mode
  DBCSR_B200_BF16_PLAN=$1 DBCSR_B200_BF16_A_TMEM=$2 timeout 400 python bench.py --config cfg4 --steps 5 --warmup 3 --no-e2e --no-cpu \
    > gpurun_out/bench_cfg4_v3_p$1_t$2.json 2> gpurun_out/bench_cfg4_v3_p$1_t$2.err
  tail -c 300 gpurun_out/bench_cfg4_v3_p$1_t$2.err
  python - "$1" "$2" <<'P'
import json, sys
for line in open("gpurun_out/bench_cfg4_v3_p%s_t%s.json" % (sys.argv[1], sys.argv[2])):
    if line.startswith("{"):
        d = json.loads(line)
        print("plan", sys.argv[1], "a_tmem", sys.argv[2], "value", d["value"], "ms", d["ms_per_step"], "selfcheck", d.get("selfcheck"))
P
done
timeout 500 python bench.py --steps 10 --warmup 3 --no-extra --no-e2e --no-cpu --no-gpu-baseline > gpurun_out/bench_r02_call12.json 2> gpurun_out/bench_r02_call12.err
tail -c 600 gpurun_out/bench_r02_call12.err
python - <<'P'
import json
for line in open("gpurun_out/bench_r02_call12.json"):
    if line.startswith("{"):
        d = json.loads(line); r = d["roofline"]
        print("value", d["value"], "kernel_only", r["kernel_only_gflops"], "burst", r["burst"]["kernel_only_gflops"], "series", r["drain_series_after_idle_ms"])
        print("tile_order", json.dumps(d.get("tile_order")))
P
