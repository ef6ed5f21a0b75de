#!/bin/bash
# round 2, call 11: planned BF16 kernel with the packed B ring -- parity tests, cfg4 per mode; FP64 bench with the fresh-operand probe
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16_tiled.py -x -q 2>&1 | tail -15 | tee gpurun_out/call11_tests.log
for mode in "1 0" "1 1"; do
  set -- $mode
  DBCSR_B200_BF16_PLAN=$1 DBCSR_B200_BF16_A_TMEM=$2 timeout 400 python bench.py --config cfg4 --steps 5 --warmup 3 --no-e2e --no-cpu \
    > gpurun_out/bench_cfg4_ring_p$1_t$2.json 2> gpurun_out/bench_cfg4_ring_p$1_t$2.err
  tail -c 300 gpurun_out/bench_cfg4_ring_p$1_t$2.err
  python - "$1" "$2" <<'P'
import json, sys
for line in open("gpurun_out/bench_cfg4_ring_p%s_t%s.json" % (sys.argv[1], sys.argv[2])):
    if line.startswith("{"):
        d = json.loads(line)
        print("plan", sys.argv[1], "a_tmem", sys.argv[2], "value", d["value"], "ms", d["ms_per_step"], "selfcheck", d.get("selfcheck"))
P
done
timeout 400 python bench.py --steps 10 --warmup 3 --no-extra --no-e2e --no-cpu --no-gpu-baseline > gpurun_out/bench_r02_call11.json 2> gpurun_out/bench_r02_call11.err
tail -c 300 gpurun_out/bench_r02_call11.err
python - <<'P'
import json
for line in open("gpurun_out/bench_r02_call11.json"):
    if line.startswith("{"):
        d = json.loads(line); r = d["roofline"]
        print("value", d["value"], "frac", r["frac"], "burst", r["burst"], "fresh", r.get("dmma_fresh_operands"), "dgemm", r.get("cublas_dgemm_8192_gflops"), r.get("cublas_dgemm_8192_sustained_gflops"))
P
