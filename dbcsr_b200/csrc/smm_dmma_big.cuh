// dbcsr_b200/csrc/smm_dmma_big.cuh -- cooperative DMMA stack kernel for blocks with a dimension in 33..80 (run-time m, n, k).
//
// The warp-autonomous kernels (smm_dmma.cuh, smm_dmma_rt.cuh) give one warp a whole block product; beyond 32x32 the accumulators
// no longer fit one warp's registers.  Here a CTA of 8 warps works on one stack entry at a time: thread 0 stages the A and B
// blocks with two cp.async.bulk copies (same 16-byte window trick as stage_block), the 8 warps form a 4 x 2 grid over the 8x8 C
// tiles (warp (wr, wc) owns tile rows wr, wr+4, wr+8 and tile columns wc, wc+2, ..., wc+8: up to 3 x 5 tiles = 30 accumulator
// registers per lane, enough for 96 x 80), every k-step each warp loads <= 3 A and <= 5 B fragments from shared memory and issues
// <= 15 DMMA.8x8x4, and a run of equal c_first stays in registers until it is flushed with RED.ADD.F64.
// Default for these shapes since its first device run (round 2: exact on integer data for all six test shapes incl. runs,
// unsorted stacks and odd alignment); libsmm_acc_b200_set_tunable("bigdmma", 0) / DBCSR_B200_BIGDMMA=0 selects the scalar generic
// kernel (smm_generic.cuh) instead.  Replaces the reference's medium/large
// kernels for this size range (src/acc/libsmm_acc/kernels/smm_acc_dnt_{medium,largeDB1,largeDB2}.h).
#pragma once
#include "smm_dmma_rt.cuh"

namespace smm {

constexpr int BIG_WARPS = 8;
constexpr int BIG_TI = 3, BIG_TJ = 5;  // tiles per warp in the 4 x 2 warp grid
constexpr int BIG_MAX_M = 96, BIG_MAX_N = 80;

__host__ __device__ inline int big_smem_bytes(int m, int n, int k) { return 128 + rt_abuf(m, k) + rt_abuf(n, k); }

__global__ void __launch_bounds__(BIG_WARPS * 32) smm_dmma_big_kernel(const int* __restrict__ stack, int stack_size,
                                                                      const double* __restrict__ a_data, const double* __restrict__ b_data,
                                                                      double* __restrict__ c_data, unsigned long long a_limit,
                                                                      unsigned long long b_limit, int chunk, int M, int N, int K) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wr = warp & 3, wc = warp >> 2;
  const int e0 = min(blockIdx.x * chunk, stack_size);
  const int e1 = min(e0 + chunk, stack_size);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");  // the predecessor may have produced A, B, C or the stack: complete + visible first
  if (e0 >= e1) {  // whole CTA
    asm volatile("griddepcontrol.wait;" ::: "memory");
    return;
  }
  const int abuf = rt_abuf(M, K);
  const uint32_t a_bytes = (uint32_t)(M * K * 8), b_bytes = (uint32_t)(N * K * 8);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  unsigned char* stg = smem_raw + 128;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int tiles_m = (M + 7) >> 3, tiles_n = (N + 7) >> 3, ksteps = (K + 3) >> 2;
  double acc[BIG_TI][BIG_TJ][2];
#pragma unroll
  for (int i = 0; i < BIG_TI; ++i)
#pragma unroll
    for (int j = 0; j < BIG_TJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto flush = [&](int c_first) {
    double* __restrict__ cb = c_data + (c_first - 1);
#pragma unroll
    for (int i = 0; i < BIG_TI; ++i) {
      const int row = (wr + 4 * i) * 8 + g;
#pragma unroll
      for (int j = 0; j < BIG_TJ; ++j) {
        const int col = (wc + 2 * j) * 8 + 2 * t;
        if (row < M) {
          if (col < N) atomicAdd(cb + (size_t)col * M + row, acc[i][j][0]);
          if (col + 1 < N) atomicAdd(cb + (size_t)(col + 1) * M + row, acc[i][j][1]);
        }
        acc[i][j][0] = acc[i][j][1] = 0.0;
      }
    }
  };

  int cur_c = -1;
  for (int e = e0; e < e1; ++e) {
    const int pa = __ldg(stack + 3 * (size_t)e), pb = __ldg(stack + 3 * (size_t)e + 1), pc = __ldg(stack + 3 * (size_t)e + 2);
    const uint64_t ga = reinterpret_cast<uint64_t>(a_data + (pa - 1));
    const uint64_t gb = reinterpret_cast<uint64_t>(b_data + (pb - 1));
    if (threadIdx.x == 0) {  // the stage was released by the __syncthreads at the end of the previous iteration
      const uint32_t ba = stage_block<0>(stg, ga, a_bytes, a_limit, bar, false, 0ull);
      const uint32_t bb = stage_block<0>(stg + abuf, gb, b_bytes, b_limit, bar, false, 0ull);
      mbar_expect_tx(bar, ba + bb);
      stage_block<0>(stg, ga, a_bytes, a_limit, bar, true, 0ull);
      stage_block<0>(stg + abuf, gb, b_bytes, b_limit, bar, true, 0ull);
    }
    if (pc != cur_c) {
      if (cur_c >= 0) flush(cur_c);
      cur_c = pc;
    }
    const double* __restrict__ As = reinterpret_cast<const double*>(stg + (uint32_t)(ga & 15ull));
    const double* __restrict__ Bs = reinterpret_cast<const double*>(stg + abuf + (uint32_t)(gb & 15ull));
    mbar_wait(bar, (uint32_t)((e - e0) & 1));
    for (int s = 0; s < ksteps; ++s) {
      const int k = 4 * s + t;
      const bool valid = k < K;
      double af[BIG_TI], bf[BIG_TJ];
#pragma unroll
      for (int i = 0; i < BIG_TI; ++i) {
        const int ti = wr + 4 * i;  // warp-uniform
        af[i] = (valid && ti < tiles_m) ? As[k * M + ti * 8 + g] : 0.0;
      }
#pragma unroll
      for (int j = 0; j < BIG_TJ; ++j) {
        const int tj = wc + 2 * j;
        bf[j] = (valid && tj < tiles_n) ? Bs[k * N + tj * 8 + g] : 0.0;
      }
#pragma unroll
      for (int i = 0; i < BIG_TI; ++i) {
        if (wr + 4 * i < tiles_m) {
#pragma unroll
          for (int j = 0; j < BIG_TJ; ++j)
            if (wc + 2 * j < tiles_n) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
      }
    }
    __syncthreads();  // every warp is done with the stage before thread 0 refills it
  }
  if (cur_c >= 0) flush(cur_c);
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

}  // namespace smm
