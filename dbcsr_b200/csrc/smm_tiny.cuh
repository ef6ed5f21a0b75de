// dbcsr_b200/csrc/smm_tiny.cuh -- FP64 stack drain for TINY blocks (m*n <= 96: the 5x5, 5x13, 13x5 families of the CP2K size set).
//
// For these shapes the tensor-pipe kernel (smm_dmma.cuh) is bound by per-entry instruction issue, not by DMMA or bandwidth: an
// entry is ~150-200 warp instructions of staging bookkeeping (TMA window arithmetic, mbarrier, shuffles) around two DMMAs that
// are 76 % padding (5 -> 8).  Here a lane OWNS C elements instead (element idx = lane, lane + 32, ...), reads its A row / B row
// values straight from global memory through L1 (a 5x5 block is two 128-byte lines, shared by the whole warp) and does K DFMAs
// per element: ~2K+10 instructions per entry, no shared memory, so 48-64 resident warps per SM hide the load latency.
// Replaces the reference's `tiny` kernel (src/acc/libsmm_acc/kernels/smm_acc_dnt_tiny.h) for the same size class; same stack
// semantics as every other drain here: C-sorted runs accumulate in registers, one RED.ADD.F64 per element and run.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "smm_dmma.cuh"

namespace smm {

constexpr int TINY_MAX_MN = 96;

template <int M, int N, int K, int WPC>
__global__ void __launch_bounds__(WPC * 32) smm_tiny_kernel(const int* __restrict__ stack, int stack_size, const double* __restrict__ a_data,
                                                            const double* __restrict__ b_data, double* __restrict__ c_data, int chunk,
                                                            int flags) {
  constexpr int MN = M * N, EPL = (MN + 31) / 32;
  static_assert(MN <= TINY_MAX_MN, "tiny kernel: at most three C elements per lane");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * WPC + warp;
  const int e0 = min(gw * chunk, stack_size), e1 = min(e0 + chunk, stack_size);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if ((flags & FLAG_PDL_CHAIN) == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (e0 >= e1) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    return;
  }
  int aoff[EPL], boff[EPL];  // element (row, col) of this lane: offsets of A(row, 0) and Bt(col, 0)
  bool act[EPL];
  double acc[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    const int idx = lane + 32 * i;
    act[i] = idx < MN;
    const int col = act[i] ? idx / M : 0;
    aoff[i] = act[i] ? idx - col * M : 0;
    boff[i] = col;
    acc[i] = 0.0;
  }
  auto flush = [&](int c_first) {
    double* __restrict__ cb = c_data + (c_first - 1);
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      if (act[i]) atomicAdd(cb + lane + 32 * i, acc[i]);
      acc[i] = 0.0;
    }
  };
  // stack entries: 32 at a time, one per lane, handed out by shuffles (no dependent global load per entry)
  int ebase = e0;
  int3 cur = make_int3(1, 1, 1);
  if (ebase + lane < e1) cur = ld_entry(stack, ebase + lane);
  int cur_c = -1;
  for (int e = e0; e < e1; ++e) {
    if (e - ebase >= 32) {
      ebase += 32;
      cur = make_int3(1, 1, 1);
      if (ebase + lane < e1) cur = ld_entry(stack, ebase + lane);
    }
    const int r = e - ebase;
    const int pa = __shfl_sync(0xffffffffu, cur.x, r), pb = __shfl_sync(0xffffffffu, cur.y, r), pc = __shfl_sync(0xffffffffu, cur.z, r);
    if (pc != cur_c) {
      if (cur_c >= 0) flush(cur_c);
      cur_c = pc;
    }
    const double* __restrict__ A = a_data + (pa - 1);
    const double* __restrict__ B = b_data + (pb - 1);
    // K is walked in slabs of KB: all loads of a slab are independent and issued back to back, then its DFMAs
    constexpr int KB = K < 8 ? K : 8;
#pragma unroll
    for (int k0 = 0; k0 < K; k0 += KB) {
      double av[EPL][KB], bv[EPL][KB];
#pragma unroll
      for (int i = 0; i < EPL; ++i)
#pragma unroll
        for (int kk = 0; kk < KB; ++kk)
          if (k0 + kk < K) {
            av[i][kk] = __ldg(A + (k0 + kk) * M + aoff[i]);
            bv[i][kk] = __ldg(B + (k0 + kk) * N + boff[i]);
          }
#pragma unroll
      for (int i = 0; i < EPL; ++i)
#pragma unroll
        for (int kk = 0; kk < KB; ++kk)
          if (k0 + kk < K) acc[i] = fma(av[i][kk], bv[i][kk], acc[i]);
    }
  }
  if (cur_c >= 0) flush(cur_c);
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

}  // namespace smm
