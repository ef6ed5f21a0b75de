#!/bin/bash
# Autotune of ALL 125 (m,n,k) triplets of the CP2K size set {5,13,23,26,32} on one GPU (about 3-4 minutes): per triplet the
# warp-autonomous DMMA kernel with warps per CTA {2,4,8} x flush {per-element RED, TMA bulk reduction} x chunking {one wave,
# run-aligned, run-aligned 12-entry chunks}, and for m*n <= 96 the lane-per-element kernel (smm_tiny.cuh) with 8 / 4 warps per CTA.
# Workload: tools/kbench.c (cfg2 structure: NBLK x NBLK block grid, 10 % occupation, 30000-entry C-sorted stacks), every run
# parity-checked (exact, integer-valued operands) against the first spec and a host recomputation.  Needs the experiment library:
#   make -C dbcsr_b200/csrc LIBNAME=libvar_exp.so BUILD=../../build/var_exp EXTRA_NVFLAGS=-DSMM_EXPERIMENT
# Output: gpurun_out/kbench_results.txt -> tools/autotune_db.py writes dbcsr_b200/parameters/parameters_B200.json from it.
mkdir -p gpurun_out
rm -f gpurun_out/kbench_results.txt gpurun_out/autotune_all.log
L=dbcsr_b200/lib/libvar_exp.so
NBLK=${NBLK:-600}
SIZES=${SIZES:-"5 13 23 26 32"}
for m in $SIZES; do for n in $SIZES; do for k in $SIZES; do
  specs="9:0:0"
  for v in 100 102 110 112 120 122; do specs="$specs $v:0:0 $v:2:0 $v:2:12"; done
  if [ $((m * n)) -le 96 ]; then specs="$specs 200:0:0 201:0:0 200:0:16 200:0:32"; fi
  timeout 60 ./tools/kbench $L gpurun_out $NBLK 0.1 2 $m,$n,$k $specs >> gpurun_out/autotune_all.log 2>&1 || echo "FAILED $m,$n,$k" >> gpurun_out/autotune_all.log
done; done; done
grep -c "parity exact" gpurun_out/kbench_results.txt
grep -c "MISMATCH" gpurun_out/kbench_results.txt
