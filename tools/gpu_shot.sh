#!/bin/bash
# One short single-GPU call: kernel-variant sweep (tools/kbench.c, pure C) on the cfg2-like workload.
mkdir -p gpurun_out
rm -f gpurun_out/kbench_results.txt
L=dbcsr_b200/lib/libvar_exp.so
timeout 25 ./tools/kbench $L gpurun_out 1000 0.1 3 23 0:0:0 60:2:10 60:2:12 60:2:14 60:2:16 60:2:20 60:0:12 61:2:12 61:2:16 60:3:12 0:2:12 > gpurun_out/kbench_23.log 2>&1
tail -12 gpurun_out/kbench_23.log
