#!/bin/bash
# round 2, call 18: panel DMMA kernel for blocks above 80 -- parity (both kernels), rate on 100^3 blocks in pure C
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_smm.py -x -q -k "large_blocks or generic or inhomogeneous" 2>&1 | tail -6 | tee gpurun_out/call18_tests.log
gcc -O2 -o tools/kbench tools/kbench.c -ldl 2>/dev/null
for h in 1 0; do
  echo "HUGEDMMA=$h"
  DBCSR_B200_HUGEDMMA=$h timeout 200 ./tools/kbench dbcsr_b200/lib/libdbcsr_acc_b200.so gpurun_out 150 0.1 2 100 0:-1:-1 2>&1 | tail -3
  DBCSR_B200_HUGEDMMA=$h timeout 200 ./tools/kbench dbcsr_b200/lib/libdbcsr_acc_b200.so gpurun_out 150 0.1 2 128,96,160 0:-1:-1 2>&1 | tail -2
done | tee gpurun_out/kbench_huge.txt
