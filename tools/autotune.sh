#!/bin/bash
# Per-shape autotune sweep of the FP64 stack kernel on one GPU (about 1 minute): for every cubic CP2K block size, warps per CTA
# {2,4,8,12,16} x flush {per-element RED, TMA bulk reduction from the operand stage} x chunking {one wave, 12 entries per warp,
# run-aligned}, on the cfg2-like 1e7-product workload of tools/kbench.c.  Needs the experiment library:
#   make -C dbcsr_b200/csrc LIBNAME=libvar_exp.so BUILD=../../build/var_exp EXTRA_NVFLAGS=-DSMM_EXPERIMENT
# Results: gpurun_out/kbench_results.txt (one line per run, parity against variant 9 = RED kernel, one wave); pick the best line
# per shape with tools/autotune_pick.py and enter it into Policy<M,N,K> (dbcsr_b200/csrc/smm_inst.cu).
mkdir -p gpurun_out
rm -f gpurun_out/kbench_results.txt
L=dbcsr_b200/lib/libvar_exp.so
for b in ${SHAPES:-23 32 26 13 5}; do
  specs="9:0:0"
  for v in 100 102 110 112 120 122 130 132 140 142; do
    specs="$specs $v:0:0 $v:2:0 $v:2:12"
  done
  if [ "$b" -le 13 ]; then  # small blocks: deeper per-warp rings (2 / 4 stages) and small chunks
    for v in 154 155 158 159; do specs="$specs $v:0:0 $v:0:6"; done
    specs="$specs 110:0:6 120:0:6"
  fi
  timeout 120 ./tools/kbench $L gpurun_out 1000 0.1 3 $b $specs > gpurun_out/kbench_$b.log 2>&1
  grep -c "parity exact" gpurun_out/kbench_$b.log
done
python tools/autotune_pick.py gpurun_out/kbench_results.txt
