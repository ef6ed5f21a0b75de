#!/bin/bash
# One single-GPU call that produces the round's evidence: GPU test suite, smoke, ncu launch list + full captures, bench lines.
set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_list_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:smm_dmma -s 340 -c 2 -o gpurun_out/prof_r01b python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:smm_bf16 -s 540 -c 2 -o gpurun_out/prof_r01_bf16 python bench.py --config cfg4 --nblk 400 --steps 1 --warmup 3 > gpurun_out/ncu_full_bf16.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err; tail -c 400 gpurun_out/bench_r01_final.json
timeout 200 python bench.py --config cfg4 --nblk 400 --steps 3 --warmup 3 > gpurun_out/bench_r01_cfg4_n400b.json 2>/dev/null; tail -c 700 gpurun_out/bench_r01_cfg4_n400b.json
timeout 120 python tools/quick_bench.py 2>&1 | tail -6
ls -la gpurun_out | tail -12
