#!/bin/bash
# round 2, call 16: ncu --set full of the planned BF16 kernel (two issuers)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:smm_bf16_planned -s 3 -c 1 -o gpurun_out/prof_r02_bf16_planned \
  python bench.py --config cfg4 --nblk 320 --steps 1 --warmup 3 --no-selfcheck --no-e2e --no-cpu > gpurun_out/ncu_bf16_planned.log 2>&1
tail -3 gpurun_out/ncu_bf16_planned.log | cut -c1-400
ls -la gpurun_out/prof_r02_bf16_planned.ncu-rep
