// dbcsr_b200/csrc/smm_dmma_rt.cuh -- DMMA stack kernel for block shapes WITHOUT a per-(m,n,k) specialisation:
// same structure as smm_dmma.cuh (warp-autonomous, TMA bulk staging of the raw blocks, DMMA.8x8x4 fragments from shared
// memory, RED flush), but m, n, k are run-time values; only the tile counts TM = ceil(m/8), TN = ceil(n/8) are compile-time
// (they size the register accumulators).  16 instantiations cover every m, n <= 32; k is bounded by the shared-memory stage.
// The k index is not permuted here (k = 4*s + t), so fragment loads can have bank conflicts; this is the "untuned" path
// (libsmm_acc_process returns 10 for it) that replaces the scalar generic kernel for small blocks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "smm_dmma.cuh"

namespace smm {

constexpr int RT_WPC = 4;  // warps per CTA, one stage each

__host__ __device__ inline int rt_abuf(int m, int k) { return (8 + (m * k + 8) * 8 + 127) & ~127; }

template <int TM, int TN>
__global__ void __launch_bounds__(RT_WPC * 32) smm_dmma_rt_kernel(const int* __restrict__ stack, int stack_size,
                                                                   const double* __restrict__ a_data, const double* __restrict__ b_data,
                                                                   double* __restrict__ c_data, unsigned long long a_limit,
                                                                   unsigned long long b_limit, int chunk, int M, int N, int K) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int gw = blockIdx.x * RT_WPC + warp;
  const int e0 = gw * chunk;
  const int e1 = min(e0 + chunk, stack_size);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");  // the predecessor may have produced A, B, C or the stack: complete + visible first
  if (e0 >= e1) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    return;
  }
  const int abuf = rt_abuf(M, K), bbuf = rt_abuf(N, K);
  const uint32_t a_bytes = (uint32_t)(M * K * 8), b_bytes = (uint32_t)(N * K * 8);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw) + warp;
  unsigned char* stg = smem_raw + 128 + (size_t)warp * (abuf + bbuf);
  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  int ebase = e0;
  int3 cur = make_int3(1, 1, 1), nxt = make_int3(1, 1, 1);
  if (ebase + lane < e1) cur = ld_entry(stack, ebase + lane);
  if (ebase + 32 + lane < e1) nxt = ld_entry(stack, ebase + 32 + lane);

  double acc[TM][TN][2];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto flush = [&](int c_first) {
    double* __restrict__ cb = c_data + (c_first - 1);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int row = i * 8 + g;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int col = j * 8 + 2 * t;
        if (row < M) {
          if (col < N) atomicAdd(cb + col * M + row, acc[i][j][0]);
          if (col + 1 < N) atomicAdd(cb + (col + 1) * M + row, acc[i][j][1]);
        }
        acc[i][j][0] = acc[i][j][1] = 0.0;
      }
    }
  };

  const int ksteps = (K + 3) >> 2;
  int cur_c = -1;
  for (int e = e0; e < e1; ++e) {
    if (e - ebase >= 32) {
      ebase += 32;
      cur = nxt;
      nxt = make_int3(1, 1, 1);
      if (ebase + 32 + lane < e1) nxt = ld_entry(stack, ebase + 32 + lane);
    }
    const int r = e - ebase;
    const int3 p = make_int3(__shfl_sync(0xffffffffu, cur.x, r), __shfl_sync(0xffffffffu, cur.y, r), __shfl_sync(0xffffffffu, cur.z, r));
    const uint64_t ga = reinterpret_cast<uint64_t>(a_data + (p.x - 1));
    const uint64_t gb = reinterpret_cast<uint64_t>(b_data + (p.y - 1));
    if (lane == 0) {  // the single stage was released by the __syncwarp at the end of the previous iteration
      const uint32_t ba = stage_block<0>(stg, ga, a_bytes, a_limit, bar, false, 0ull);
      const uint32_t bb = stage_block<0>(stg + abuf, gb, b_bytes, b_limit, bar, false, 0ull);
      mbar_expect_tx(bar, ba + bb);
      stage_block<0>(stg, ga, a_bytes, a_limit, bar, true, 0ull);
      stage_block<0>(stg + abuf, gb, b_bytes, b_limit, bar, true, 0ull);
    }
    if (p.z != cur_c) {
      if (cur_c >= 0) flush(cur_c);
      cur_c = p.z;
    }
    const double* __restrict__ As = reinterpret_cast<const double*>(stg + (uint32_t)(ga & 15ull));
    const double* __restrict__ Bs = reinterpret_cast<const double*>(stg + abuf + (uint32_t)(gb & 15ull));
    mbar_wait(bar, (uint32_t)((e - e0) & 1));
    for (int s = 0; s < ksteps; ++s) {
      const int k = 4 * s + t;
      const bool valid = k < K;
      double af[TM], bf[TN];
#pragma unroll
      for (int ti = 0; ti < TM; ++ti) af[ti] = valid ? As[k * M + ti * 8 + g] : 0.0;
#pragma unroll
      for (int tj = 0; tj < TN; ++tj) bf[tj] = valid ? Bs[k * N + tj * 8 + g] : 0.0;
#pragma unroll
      for (int ti = 0; ti < TM; ++ti)
#pragma unroll
        for (int tj = 0; tj < TN; ++tj) dmma884(acc[ti][tj][0], acc[ti][tj][1], af[ti], bf[tj]);
    }
    __syncwarp();
  }
  if (cur_c >= 0) flush(cur_c);
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

}  // namespace smm
