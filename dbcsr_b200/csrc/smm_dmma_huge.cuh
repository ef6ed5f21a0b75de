// dbcsr_b200/csrc/smm_dmma_huge.cuh -- DMMA stack kernel for blocks with a dimension above max_kernel_dim (80), any m, n, k.
//
// The reference loops cublasDgemm over the host stack for these blocks and synchronises (src/acc/libsmm_acc/libsmm_acc.cpp:256-278);
// round 1 drained them with the scalar generic kernel (one warp per entry, DFMA).  Here a CTA of 8 warps works on one entry at a
// time as a small blocked GEMM on the FP64 tensor pipe: C is cut into panels of <= 96 x 80 (equal parts, so that 100 x 100 becomes
// four 56 x 56 panels and not 96 + 4), K into chunks of 32; per chunk the CTA copies the A rows (column-major: contiguous runs per k)
// and the B columns of the panel into padded shared-memory tiles As[kk][row], Bs[kk][col] (leading dimensions = 4 mod 16 doubles:
// the 8 x 4 fragment pattern of DMMA.8x8x4 then needs exactly its two wavefronts), the 8 warps form the 4 x 2 grid over the 8 x 8
// C tiles of smm_dmma_big.cuh (<= 3 x 5 tiles = 30 accumulators per lane) and issue DMMA.8x8x4; a finished panel is added to C
// with RED.ADD.F64.  B blocks arrive transposed (n x k) only when both n and k fit max_kernel_dim (libsmm_acc.cpp:267-270): both
// layouts are read with coalesced global loads (along n when transposed, along k when not).
#pragma once
#include "smm_dmma_big.cuh"

namespace smm {

constexpr int HUGE_KC = 32;
constexpr int HUGE_LDA = BIG_MAX_M + 4;   // 100: 4 mod 16
constexpr int HUGE_LDB = BIG_MAX_N + 4;   // 84: 4 mod 16
constexpr int HUGE_SMEM = (HUGE_KC * HUGE_LDA + HUGE_KC * HUGE_LDB) * 8;  // 47104 B: no opt-in needed

__global__ void __launch_bounds__(BIG_WARPS * 32) smm_dmma_huge_kernel(const int* __restrict__ stack, int stack_size,
                                                                       const double* __restrict__ a_data, const double* __restrict__ b_data,
                                                                       double* __restrict__ c_data, int M, int N, int K, int b_transposed) {
  __shared__ __align__(16) double As[HUGE_KC * HUGE_LDA];
  __shared__ __align__(16) double Bs[HUGE_KC * HUGE_LDB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int g = lane >> 2, t = lane & 3;
  const int wr = warp & 3, wc = warp >> 2;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");  // the predecessor may have produced A, B, C or the stack
  // equal panels: pm rows, pn columns (multiples of 8 except at the block edge)
  const int npm = (M + BIG_MAX_M - 1) / BIG_MAX_M, npn = (N + BIG_MAX_N - 1) / BIG_MAX_N;
  const int pm = (((M + npm - 1) / npm) + 7) & ~7, pn = (((N + npn - 1) / npn) + 7) & ~7;
  for (int e = blockIdx.x; e < stack_size; e += gridDim.x) {
    const int pa = __ldg(stack + 3 * (size_t)e), pb = __ldg(stack + 3 * (size_t)e + 1), pc = __ldg(stack + 3 * (size_t)e + 2);
    const double* __restrict__ A = a_data + (pa - 1);
    const double* __restrict__ B = b_data + (pb - 1);
    double* __restrict__ C = c_data + (pc - 1);
    for (int m0 = 0; m0 < M; m0 += pm) {
      const int mrows = min(pm, M - m0);
      const int tiles_m = (mrows + 7) >> 3;
      for (int n0 = 0; n0 < N; n0 += pn) {
        const int ncols = min(pn, N - n0);
        const int tiles_n = (ncols + 7) >> 3;
        double acc[BIG_TI][BIG_TJ][2];
#pragma unroll
        for (int i = 0; i < BIG_TI; ++i)
#pragma unroll
          for (int j = 0; j < BIG_TJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int k0 = 0; k0 < K; k0 += HUGE_KC) {
          const int kc = min(HUGE_KC, K - k0);
          // A panel chunk: (row, kk) <- A[(k0 + kk) * M + m0 + row]; rows beyond the panel and k beyond the block read as zeros
          for (int i = tid; i < HUGE_KC * (tiles_m * 8); i += BIG_WARPS * 32) {
            const int kk = i / (tiles_m * 8), r = i - kk * (tiles_m * 8);
            As[kk * HUGE_LDA + r] = (kk < kc && r < mrows) ? __ldg(A + (size_t)(k0 + kk) * M + m0 + r) : 0.0;
          }
          if (b_transposed) {  // Bt is n x k column-major: (col, kk) at B[(k0 + kk) * N + n0 + col]
            for (int i = tid; i < HUGE_KC * (tiles_n * 8); i += BIG_WARPS * 32) {
              const int kk = i / (tiles_n * 8), cidx = i - kk * (tiles_n * 8);
              Bs[kk * HUGE_LDB + cidx] = (kk < kc && cidx < ncols) ? __ldg(B + (size_t)(k0 + kk) * N + n0 + cidx) : 0.0;
            }
          }
          else {  // B is k x n column-major: (kk, col) at B[(n0 + col) * K + k0 + kk]: consecutive threads walk k
            for (int i = tid; i < HUGE_KC * (tiles_n * 8); i += BIG_WARPS * 32) {
              const int cidx = i / HUGE_KC, kk = i - cidx * HUGE_KC;
              Bs[kk * HUGE_LDB + cidx] = (kk < kc && cidx < ncols) ? __ldg(B + (size_t)(n0 + cidx) * K + k0 + kk) : 0.0;
            }
          }
          __syncthreads();
#pragma unroll 2
          for (int s = 0; s < HUGE_KC / 4; ++s) {
            const int kk = 4 * s + t;
            double af[BIG_TI], bf[BIG_TJ];
#pragma unroll
            for (int i = 0; i < BIG_TI; ++i) {
              const int ti = wr + 4 * i;  // warp-uniform
              af[i] = ti < tiles_m ? As[kk * HUGE_LDA + ti * 8 + g] : 0.0;
            }
#pragma unroll
            for (int j = 0; j < BIG_TJ; ++j) {
              const int tj = wc + 2 * j;
              bf[j] = tj < tiles_n ? Bs[kk * HUGE_LDB + tj * 8 + g] : 0.0;
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
          __syncthreads();  // every warp is done with the tiles before they are refilled
        }
#pragma unroll
        for (int i = 0; i < BIG_TI; ++i) {
          const int row = (wr + 4 * i) * 8 + g;
#pragma unroll
          for (int j = 0; j < BIG_TJ; ++j) {
            const int col = (wc + 2 * j) * 8 + 2 * t;
            if (row < mrows) {
              if (col < ncols) atomicAdd(C + (size_t)(n0 + col) * M + m0 + row, acc[i][j][0]);
              if (col + 1 < ncols) atomicAdd(C + (size_t)(n0 + col + 1) * M + m0 + row, acc[i][j][1]);
            }
          }
        }
      }
    }
  }
}

}  // namespace smm
