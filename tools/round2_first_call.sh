#!/bin/bash
# First single-GPU call of round 2 (about 10 GPU-minutes): what round 1 could not measure any more.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/round2_first_call.sh'
# Needs (built here, they travel with the snapshot): the production library, dbcsr_b200/lib/libvar_exp.so
#   (make -C dbcsr_b200/csrc LIBNAME=libvar_exp.so BUILD=../../build/var_exp EXTRA_NVFLAGS=-DSMM_EXPERIMENT) and tools/kbench.
set -x
mkdir -p gpurun_out
P=dbcsr_b200/lib/libdbcsr_acc_b200.so
X=dbcsr_b200/lib/libvar_exp.so
# 1. shipped policy vs the RED kernel, parity + kernel-only TFLOP/s (production library: variant field is ignored, 0:0:0 = RED-less
#    one-wave split of the shipped kernel, 0:-1:-1 = shipped policy); then the experiment library with the unmeasured variants
timeout 60 ./tools/kbench $P gpurun_out 1000 0.1 3 23 0:-1:-1 0:0:0 0:2:0 0:2:12 > gpurun_out/kbench_prod.log 2>&1; tail -6 gpurun_out/kbench_prod.log
timeout 60 ./tools/kbench $X gpurun_out 1000 0.1 3 23 9:0:0 0:-1:-1 70:2:12 71:2:12 75:2:12 76:2:12 70:0:0 75:0:0 72:2:12:t > gpurun_out/kbench_new_variants.log 2>&1; tail -10 gpurun_out/kbench_new_variants.log
# 2. per-shape autotune sweep
timeout 300 bash tools/autotune.sh > gpurun_out/autotune.log 2>&1; tail -60 gpurun_out/autotune.log
# 3. ncu: launch list + full capture of the shipped 23^3 kernel (no Python start-up)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_ncu_launches.csv \
  ./tools/kbench $P gpurun_out 1000 0.1 1 23 0:-1:-1 > gpurun_out/ncu_list.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:smm_dmma -s 400 -c 2 -o gpurun_out/prof_r02 \
  ./tools/kbench $P gpurun_out 1000 0.1 1 23 0:-1:-1 > gpurun_out/ncu_full.log 2>&1
# 4. GPU test suite, smoke, bench
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; tail -c 1500 gpurun_out/bench_r02_n1.json
# 5. the opt-in cooperative kernel for 33..80 blocks (first device run; under its own timeout)
DBCSR_B200_TEST_UNVERIFIED=1 timeout 120 python -m pytest tests/test_gpu_zz_dbcsr_multiply.py -q -m gpu -k cooperative 2>&1 | tail -5
ls -la gpurun_out | tail -15
