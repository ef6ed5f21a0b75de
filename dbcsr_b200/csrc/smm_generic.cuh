// dbcsr_b200/csrc/smm_generic.cuh -- shape-agnostic kernels: generic stack drain (any m,n,k; B transposed or not),
// batched in-place block transpose, batched block norms.
//
// Reference counterparts: the untuned default of src/acc/libsmm_acc/libsmm_acc.cpp:222-231 and the per-entry cuBLAS loop
// :256-278 (generic drain); src/acc/libsmm_acc/kernels/smm_acc_transpose.h:41-65 (transpose);
// src/acc/cuda_hip/calculate_norms.cpp:48-117 (norms).  All new code; run-time shapes, no JIT.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "smm_dmma.cuh"

namespace smm {

// ---- element types of the ABI (libsmm_acc_data_t): real_4, real_8, complex_4, complex_8 ---------------------------------------
template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static __device__ __forceinline__ float zero() { return 0.f; }
  static __device__ __forceinline__ void fma_(float& acc, float a, float b) { acc = fmaf(a, b, acc); }
  static __device__ __forceinline__ void add(float& acc, float v) { acc += v; }
  static __device__ __forceinline__ void red(float* p, float v) { atomicAdd(p, v); }
};
template <>
struct Elem<double> {
  static __device__ __forceinline__ double zero() { return 0.0; }
  static __device__ __forceinline__ void fma_(double& acc, double a, double b) { acc = fma(a, b, acc); }
  static __device__ __forceinline__ void add(double& acc, double v) { acc += v; }
  static __device__ __forceinline__ void red(double* p, double v) { atomicAdd(p, v); }
};
template <>
struct Elem<float2> {
  static __device__ __forceinline__ float2 zero() { return make_float2(0.f, 0.f); }
  static __device__ __forceinline__ void fma_(float2& acc, float2 a, float2 b) {
    acc.x = fmaf(a.x, b.x, fmaf(-a.y, b.y, acc.x));
    acc.y = fmaf(a.x, b.y, fmaf(a.y, b.x, acc.y));
  }
  static __device__ __forceinline__ void add(float2& acc, float2 v) {
    acc.x += v.x;
    acc.y += v.y;
  }
  static __device__ __forceinline__ void red(float2* p, float2 v) {
    atomicAdd(&p->x, v.x);
    atomicAdd(&p->y, v.y);
  }
};
template <>
struct Elem<double2> {
  static __device__ __forceinline__ double2 zero() { return make_double2(0.0, 0.0); }
  static __device__ __forceinline__ void fma_(double2& acc, double2 a, double2 b) {
    acc.x = fma(a.x, b.x, fma(-a.y, b.y, acc.x));
    acc.y = fma(a.x, b.y, fma(a.y, b.x, acc.y));
  }
  static __device__ __forceinline__ void add(double2& acc, double2 v) {
    acc.x += v.x;
    acc.y += v.y;
  }
  static __device__ __forceinline__ void red(double2* p, double2 v) {
    atomicAdd(&p->x, v.x);
    atomicAdd(&p->y, v.y);
  }
};

// Generic drain for the non-real_8 types (the reference rejects them with -10 and DBCSR computes them on the CPU,
// src/acc/libsmm_acc/libsmm_acc.cpp:338): one warp per entry, lanes own C elements, plain C += A*B (no conjugation).
template <typename T>
__global__ void __launch_bounds__(256) smm_generic_typed_kernel(const int* __restrict__ stack, int stack_size, const T* __restrict__ a_data,
                                                               const T* __restrict__ b_data, T* __restrict__ c_data, int m, int n, int k,
                                                               int b_transposed, int chunk) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 8 + warp;
  const int e0 = gw * chunk;
  const int e1 = min(e0 + chunk, stack_size);
  const int mn = m * n;
  for (int e = e0; e < e1; ++e) {
    const int3 p = ld_entry(stack, e);
    const T* __restrict__ A = a_data + (p.x - 1);
    const T* __restrict__ B = b_data + (p.y - 1);
    T* cb = c_data + (p.z - 1);
    for (int idx = lane; idx < mn; idx += 32) {
      const int col = idx / m, row = idx - col * m;
      T s = Elem<T>::zero();
      if (b_transposed) {
        for (int l = 0; l < k; ++l) Elem<T>::fma_(s, A[l * m + row], B[l * n + col]);
      }
      else {
        for (int l = 0; l < k; ++l) Elem<T>::fma_(s, A[l * m + row], B[col * k + l]);
      }
      Elem<T>::red(cb + idx, s);
    }
  }
}

// In-place transpose for any element type (one warp per block through shared memory), see transpose_kernel below.
template <typename T>
__global__ void transpose_typed_kernel(const int* __restrict__ trs_stack, int nblks, T* __restrict__ data, int m, int n) {
  extern __shared__ unsigned char tr_raw[];
  T* tr = reinterpret_cast<T*>(tr_raw);
  const int wpc = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mn = m * n;
  T* buf = tr + (size_t)warp * mn;
  for (int b = blockIdx.x * wpc + warp; b < nblks; b += gridDim.x * wpc) {
    T* blk = data + __ldg(trs_stack + b);
    for (int i = lane; i < mn; i += 32) buf[i] = blk[i];
    __syncwarp();
    for (int i = lane; i < mn; i += 32) blk[i] = buf[(i % n) * m + i / n];
    __syncwarp();
  }
}

// One warp per stack entry (grid-stride over chunks of the C-sorted stack); lanes own C elements round-robin.
// A and B are read straight from global memory (L1/L2 serve the reuse inside a block product); consecutive entries with the
// same c_first are accumulated in registers when m*n <= 32*GEN_ACC, otherwise every entry is flushed on its own.
constexpr int GEN_ACC = 8;   // C elements per lane held in registers
constexpr int GEN_WPC = 8;   // warps per CTA

__global__ void __launch_bounds__(GEN_WPC * 32) smm_generic_kernel(const int* __restrict__ stack, int stack_size,
                                                                    const double* __restrict__ a_data, const double* __restrict__ b_data,
                                                                    double* __restrict__ c_data, int m, int n, int k, int b_transposed,
                                                                    int chunk) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * GEN_WPC + warp;
  const int e0 = gw * chunk;
  const int e1 = min(e0 + chunk, stack_size);
  if (e0 >= e1) return;
  const int mn = m * n;

  if (mn <= 32 * GEN_ACC) {
    double acc[GEN_ACC];
#pragma unroll
    for (int q = 0; q < GEN_ACC; ++q) acc[q] = 0.0;
    int cur_c = -1;
    for (int e = e0; e <= e1; ++e) {
      int3 p = make_int3(0, 0, -2);
      if (e < e1) p = ld_entry(stack, e);
      if (p.z != cur_c) {
        if (cur_c >= 0) {
          double* cb = c_data + (cur_c - 1);
#pragma unroll
          for (int q = 0; q < GEN_ACC; ++q) {
            const int idx = lane + 32 * q;
            if (idx < mn) atomicAdd(cb + idx, acc[q]);
            acc[q] = 0.0;
          }
        }
        cur_c = p.z;
      }
      if (e == e1) break;
      const double* __restrict__ A = a_data + (p.x - 1);
      const double* __restrict__ B = b_data + (p.y - 1);
#pragma unroll
      for (int q = 0; q < GEN_ACC; ++q) {
        const int idx = lane + 32 * q;
        if (idx < mn) {
          const int col = idx / m, row = idx - col * m;
          double s = 0.0;
          if (b_transposed) {
            for (int l = 0; l < k; ++l) s = fma(__ldg(A + l * m + row), __ldg(B + l * n + col), s);
          }
          else {
            for (int l = 0; l < k; ++l) s = fma(__ldg(A + l * m + row), __ldg(B + col * k + l), s);
          }
          acc[q] += s;
        }
      }
    }
  }
  else {
    for (int e = e0; e < e1; ++e) {
      const int3 p = ld_entry(stack, e);
      const double* __restrict__ A = a_data + (p.x - 1);
      const double* __restrict__ B = b_data + (p.y - 1);
      double* cb = c_data + (p.z - 1);
      for (int idx = lane; idx < mn; idx += 32) {
        const int col = idx / m, row = idx - col * m;
        double s = 0.0;
        if (b_transposed) {
          for (int l = 0; l < k; ++l) s = fma(__ldg(A + l * m + row), __ldg(B + l * n + col), s);
        }
        else {
          for (int l = 0; l < k; ++l) s = fma(__ldg(A + l * m + row), __ldg(B + col * k + l), s);
        }
        atomicAdd(cb + idx, s);
      }
    }
  }
}

// In-place transpose of m x n col-major blocks (result n x m col-major): out[i] = in[(i % n) * m + i / n].
// One warp per block, grid-stride; the block is parked in the warp's slice of dynamic shared memory.
// Fused norms (norms != nullptr): while the block sits in shared memory its sum of squares is reduced and stored as float at
// norms[trs_blk[b]] (trs_blk: position of the block in the panel's list; nullptr = b) -- what c_calculate_norms would compute in
// a second pass over the panel (src/acc/cuda_hip/calculate_norms.cpp:48-96); the transpose does not change a block's norm.
__global__ void transpose_kernel(const int* __restrict__ trs_stack, int nblks, double* __restrict__ data, int m, int n,
                                 const int* __restrict__ trs_blk, float* __restrict__ norms) {
  extern __shared__ double tr_smem[];
  const int wpc = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mn = m * n;
  double* buf = tr_smem + (size_t)warp * mn;
  for (int b = blockIdx.x * wpc + warp; b < nblks; b += gridDim.x * wpc) {
    double* blk = data + __ldg(trs_stack + b);
    double ss = 0.0;
    for (int i = lane; i < mn; i += 32) {
      const double v = blk[i];
      buf[i] = v;
      ss = fma(v, v, ss);
    }
    __syncwarp();
    for (int i = lane; i < mn; i += 32) {
      const int r_out = i % n, c_out = i / n;
      blk[i] = buf[r_out * m + c_out];
    }
    if (norms != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      if (lane == 0) norms[trs_blk != nullptr ? __ldg(trs_blk + b) : b] = (float)ss;
    }
    __syncwarp();
  }
}

// norms[b] = sum_i mat[offsets[b] + i]^2 (accumulated in double, stored as OUT = float for c_calculate_norms, double for the
// final filter); one warp per block, grid-stride.
template <typename OUT>
__global__ void norms_kernel(const double* __restrict__ mat, int nblks, const int* __restrict__ offsets, const int* __restrict__ nelems,
                             OUT* __restrict__ norms) {
  const int wpc = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = blockIdx.x * wpc + warp; b < nblks; b += gridDim.x * wpc) {
    const double* __restrict__ p = mat + __ldg(offsets + b);
    const int ne = __ldg(nelems + b);
    double s = 0.0;
    for (int i = lane; i < ne; i += 32) {
      const double d = __ldg(p + i);
      s = fma(d, d, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) norms[b] = (OUT)s;
  }
}

// dst[dst_off[b] + i] = src[src_off[b] + i], i < nelems[b]: compaction of the C blocks that survive the final filter (and any
// other block-wise re-packing of a data area).  One warp per block, grid-stride; pure HBM streaming (16 B/element).
__global__ void gather_blocks_kernel(const double* __restrict__ src, double* __restrict__ dst, int nblks, const int* __restrict__ src_off,
                                     const int* __restrict__ dst_off, const int* __restrict__ nelems) {
  const int wpc = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = blockIdx.x * wpc + warp; b < nblks; b += gridDim.x * wpc) {
    const double* __restrict__ p = src + __ldg(src_off + b);
    double* __restrict__ q = dst + __ldg(dst_off + b);
    const int ne = __ldg(nelems + b);
    for (int i = lane; i < ne; i += 32) q[i] = __ldg(p + i);
  }
}

}  // namespace smm
