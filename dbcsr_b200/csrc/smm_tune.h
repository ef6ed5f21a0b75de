// dbcsr_b200/csrc/smm_tune.h -- run-time knobs of the FP64 stack kernels, shared by the per-shape launch tables (smm_inst.cu)
// and the C ABI (libsmm_api.cu: libsmm_acc_b200_set_tunable / _get_tunable).  Defaults come from the environment once at load:
//   DBCSR_B200_BALANCE=0|1  split a stack over the warps in chunks that differ by at most one entry (default 0: measured slower, it
//                           leaves no free CTA slots for the next launch to start in)
//   DBCSR_B200_ALIGN=0|1    move chunk boundaries to changes of c_first, so that a run is flushed once (default: per-shape policy)
//   DBCSR_B200_CHUNK=n      n > 0: at most n entries per warp; the grid may then exceed one resident wave; 0 = one wave
//                           (default: per-shape policy)
//   DBCSR_B200_BIGDMMA=0|1  cooperative DMMA kernel for 33..80 blocks instead of the scalar generic kernel (default 1; verified on B200 in round 2)
//   DBCSR_B200_INHOMOGENEOUS=0|1  0 = reject inhomogeneous stacks with -1 like the reference (DBCSR then uses its CPU driver); default 1
//   DBCSR_B200_VARIANT=v    kernel variant id (see smm_inst.cu; only experiment builds carry more than the default)
// The trace window (kernel-timeline instrumentation, tools/kbench.c + tools/trace_analyze.py) is set through the setter only.
#pragma once
#include <atomic>

namespace smm {

constexpr int TRACE_WARPS = 4096;  // trace records per launch slot
constexpr int TRACE_REC = 128;     // 64-bit words per warp record

struct Tunables {
  std::atomic<int> variant{0};
  std::atomic<int> hugedmma{1};  // 1 (default): blocks with a dimension above max_kernel_dim use the panel DMMA kernel (smm_dmma_huge.cuh); 0: scalar generic kernel
  std::atomic<int> bigdmma{1};  // 1 (default): blocks with a dimension in 33..80 use the cooperative DMMA kernel (smm_dmma_big.cuh); 0: scalar generic kernel
  std::atomic<int> inhomogeneous{1};  // 1: def_mnk = 0 stacks are binned by shape and drained on the GPU; 0: -1 like the reference
  std::atomic<int> bf16_merge{1};     // tiled BF16 SpGEMM: 1 = one wide MMA per run of adjacent existing B blocks, 0 = one per block
  std::atomic<int> bf16_plan{1};      // tiled BF16 SpGEMM: 1 = planned kernel (smm_bf16_plan.cuh: copy commands / MMA runs derived once per multiply)
  std::atomic<int> bf16_a_tmem{0};    // tiled BF16 SpGEMM: 1 = A operand staged in TMEM by tcgen05.cp (15 block columns per tile)
  std::atomic<int> balance{0};
  std::atomic<int> chunk{-1};  // -1 = per-shape policy (smm_inst.cu), 0 = one resident wave, n > 0 = n entries per warp
  std::atomic<int> align{-1};  // -1 = per-shape policy, 0 / 1 = off / on
  std::atomic<unsigned long long*> trace{nullptr};  // device buffer of trace_count * TRACE_WARPS * TRACE_REC words
  std::atomic<int> trace_first{0};                  // first launch sequence number that is recorded
  std::atomic<int> trace_count{0};                  // number of consecutive launches recorded
  std::atomic<int> seq{0};                          // launch sequence number of the tuned FP64 kernels
};
extern Tunables g_tune;

}  // namespace smm
