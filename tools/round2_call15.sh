#!/bin/bash
# round 2, call 15: BF16 planned kernel v5 (A in TMEM, four issuers)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bf16_tiled.py -x -q 2>&1 | tail -8 | tee gpurun_out/call15_tests.log
for st in 5 4; do
DBCSR_B200_BF16_STAGES=$st timeout 400 python bench.py --config cfg4 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_cfg4_v5_st$st.json 2> gpurun_out/bench_cfg4_v5_st$st.err
tail -c 300 gpurun_out/bench_cfg4_v5_st$st.err
python - $st <<'P'
import json, sys
for line in open("gpurun_out/bench_cfg4_v5_st%s.json" % sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        print("stages", sys.argv[1], "cfg4 value", d["value"], "ms", d["ms_per_step"], "selfcheck", d.get("selfcheck"))
P
done
