#!/bin/bash
# Third single-GPU call of round 2: full GPU suite (fixed tiled BF16 geometry, merged-run MMAs, multi-rank Cannon parity on a shared
# GPU), chain vs waiting mode of the shipped 23^3 kernel in pure C, autotune of all 125 triplets in chain mode, bench, ncu captures.
set -x
mkdir -p gpurun_out
P=dbcsr_b200/lib/libdbcsr_acc_b200.so
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
rm -f gpurun_out/kbench_results.txt
timeout 60 ./tools/kbench $P gpurun_out 1000 0.1 3 23 0:-1:-1 0:0:0 > gpurun_out/kbench_chain_23.log 2>&1; grep -E "chain mode|TFLOP" gpurun_out/kbench_chain_23.log
KBENCH_NO_CHAIN=1 timeout 60 ./tools/kbench $P gpurun_out 1000 0.1 3 23 0:-1:-1 0:0:0 > gpurun_out/kbench_nochain_23.log 2>&1; grep -E "chain mode|TFLOP" gpurun_out/kbench_nochain_23.log
cp gpurun_out/kbench_results.txt gpurun_out/kbench_chain_vs_wait.txt
timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02_call3.json 2> gpurun_out/bench_r02_call3.err; tail -c 2500 gpurun_out/bench_r02_call3.json; tail -5 gpurun_out/bench_r02_call3.err
timeout 200 python bench.py --steps 3 --warmup 3 --no-extra --no-gpu-baseline --no-cpu --no-selfcheck --threads 16 > gpurun_out/bench_r02_call3_t16.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_call3_t16.json').read().strip().splitlines()[-1]); print('e2e 16 threads', d['e2e'])"
# ncu: tiled BF16 kernel (small grid so that the replays stay short) and the five cubic FP64 kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:smm_bf16_tiled -s 3 -c 1 -o gpurun_out/prof_r02_bf16_tiled \
  python bench.py --config cfg4 --nblk 320 --steps 1 --warmup 3 --no-selfcheck > gpurun_out/ncu_bf16.log 2>&1; tail -3 gpurun_out/ncu_bf16.log
for b in 5 13 26 32; do
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:smm_ -s 200 -c 1 -o gpurun_out/prof_r02_fp64_$b \
    ./tools/kbench $P gpurun_out 1000 0.1 1 $b 0:-1:-1 > gpurun_out/ncu_fp64_$b.log 2>&1; tail -1 gpurun_out/ncu_fp64_$b.log
done
timeout 600 bash tools/autotune_all.sh 2>&1 | tail -4
python tools/autotune_db.py gpurun_out/kbench_results.txt --dry > gpurun_out/autotune_pick.txt 2>&1; head -3 gpurun_out/autotune_pick.txt
ls -la gpurun_out | tail -12
