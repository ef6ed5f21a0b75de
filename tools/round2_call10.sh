#!/bin/bash
# round 2, call 10: planned tiled BF16 kernel -- parity tests (all modes), then cfg4 at the full grid per mode
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16_tiled.py -x -q 2>&1 | tail -15 | tee gpurun_out/call10_tests.log
for mode in "1 0" "1 1" "0 0"; do
  set -- $mode
  DBCSR_B200_BF16_PLAN=$1 DBCSR_B200_BF16_A_TMEM=$2 timeout 400 python bench.py --config cfg4 --steps 5 --warmup 3 --no-e2e --no-cpu \
    > gpurun_out/bench_cfg4_p$1_t$2.json 2> gpurun_out/bench_cfg4_p$1_t$2.err
  tail -c 300 gpurun_out/bench_cfg4_p$1_t$2.err
  python - "$1" "$2" <<'P'
import json, sys
for line in open("gpurun_out/bench_cfg4_p%s_t%s.json" % (sys.argv[1], sys.argv[2])):
    if line.startswith("{"):
        d = json.loads(line)
        print("plan", sys.argv[1], "a_tmem", sys.argv[2], "value", d["value"], "ms", d["ms_per_step"], "selfcheck", d.get("selfcheck"))
P
done
