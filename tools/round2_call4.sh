#!/bin/bash
# Fourth single-GPU call of round 2: whole GPU suite without -x (new engine / cannon / smm tests have not run yet), sustained vs
# burst rate of the shipped 23^3 kernel in pure C, tiled BF16 kernel modes, bench, ncu of the BF16 kernel.
set -x
mkdir -p gpurun_out
P=dbcsr_b200/lib/libdbcsr_acc_b200.so
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40
rm -f gpurun_out/kbench_results.txt
timeout 120 ./tools/kbench $P gpurun_out 1000 0.1 3 23 0:-1:-1 > gpurun_out/kbench_burst_23.log 2>&1; grep -E "TFLOP" gpurun_out/kbench_burst_23.log
timeout 120 ./tools/kbench $P gpurun_out 1000 0.1 80 23 0:-1:-1 > gpurun_out/kbench_sustained_23.log 2>&1; grep -E "TFLOP" gpurun_out/kbench_sustained_23.log
for mode in "1 0" "0 0" "1 1" "0 1"; do
  set -- $mode
  DBCSR_B200_BF16_MERGE=$1 DBCSR_B200_BF16_A_TMEM=$2 timeout 200 python bench.py --config cfg4 --steps 5 --warmup 3 > gpurun_out/bench_cfg4_m$1_t$2.json 2> gpurun_out/bench_cfg4_m$1_t$2.err
  python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/bench_cfg4_m$1_t$2.json').read().strip().splitlines()[-1]); print('cfg4 merge=$1 a_tmem=$2', d['value'], d['ms_per_step'], d['selfcheck'])
except Exception as ex: print('cfg4 merge=$1 a_tmem=$2 FAILED', ex); print(open('gpurun_out/bench_cfg4_m$1_t$2.err').read()[-1500:])"
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02_call4.json 2> gpurun_out/bench_r02_call4.err; tail -c 1500 gpurun_out/bench_r02_call4.json; tail -5 gpurun_out/bench_r02_call4.err
for t in 0 1; do
DBCSR_B200_BF16_A_TMEM=$t timeout 300 ncu --set full --clock-control none --import-source on -k regex:smm_bf16_tiled -s 3 -c 1 -o gpurun_out/prof_r02_bf16_tiled_t$t \
  python bench.py --config cfg4 --nblk 320 --steps 1 --warmup 3 --no-selfcheck > gpurun_out/ncu_bf16_t$t.log 2>&1; tail -2 gpurun_out/ncu_bf16_t$t.log
done
ls -la gpurun_out | tail -12
