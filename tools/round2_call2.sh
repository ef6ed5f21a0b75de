#!/bin/bash
# Second single-GPU call of round 2: full GPU test suite (new kernels / engine paths), reference-kernel GPU baseline for the five
# cubic shapes, autotune of all 125 triplets, bench with extra configs.
set -x
mkdir -p gpurun_out
P=dbcsr_b200/lib/libdbcsr_acc_b200.so
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
# reference libsmm_acc kernels (NVRTC JIT for compute_100, H100 parameter set) on this GPU, same stacks, same ABI
for b in 23 32 26 13 5; do
  KBENCH_ACC_LIB=baseline/_ref/libdbcsr_acc_ref.so timeout 120 ./tools/kbench $P gpurun_out 1000 0.1 3 $b 0:0:0 > gpurun_out/kbench_ref_$b.log 2>&1
  tail -3 gpurun_out/kbench_ref_$b.log
  timeout 60 ./tools/kbench $P gpurun_out 1000 0.1 3 $b 0:-1:-1 > gpurun_out/kbench_ours_$b.log 2>&1
  tail -1 gpurun_out/kbench_ours_$b.log
done
cp gpurun_out/kbench_results.txt gpurun_out/kbench_ref_vs_ours.txt
timeout 600 bash tools/autotune_all.sh 2>&1 | tail -4
python tools/autotune_db.py gpurun_out/kbench_results.txt --dry > gpurun_out/autotune_pick.txt 2>&1; tail -130 gpurun_out/autotune_pick.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02_call2.json 2> gpurun_out/bench_r02_call2.err; tail -c 6000 gpurun_out/bench_r02_call2.json; tail -20 gpurun_out/bench_r02_call2.err
ls -la gpurun_out | tail -12
