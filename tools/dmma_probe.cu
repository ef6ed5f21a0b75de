// tools/dmma_probe.cu -- does ONE warp per SM sub-partition reach the FP64 tensor-pipe rate when every DMMA.8x8x4 reads
// different operand registers (as the stack kernel does), or only with reused operands (tools/microbench.cu)?
// Modes: 0 = same a,b for all 9 accumulators (microbench pattern); 1 = 3 a x 3 b registers (kernel pattern, values fixed);
//        2 = like 1, and the a/b registers are re-loaded from shared memory every k-step (LDS.64, conflict-free).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void probe(double* out, int iters, double x, long long* cycles) {
  __shared__ double sm[32 * 64];
  for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) sm[i] = x + i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double c0[9], c1[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) c0[i] = c1[i] = 0;
  double a[3] = {x + lane, x + 2 * lane, x - lane}, b[3] = {x * 2, x * 3 + lane, x * 5};
  const double* p = sm + lane + (warp & 1) * 1024;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 3; ++i) { a[i] = p[((it & 7) * 6 + i) * 32 % 1024]; b[i] = p[((it & 7) * 6 + 3 + i) * 32 % 1024]; }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (MODE == 0) dmma884(c0[i * 3 + j], c1[i * 3 + j], a[0], b[0]);
        else dmma884(c0[i * 3 + j], c1[i * 3 + j], a[j], b[i]);
      }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}
template <int MODE>
int run(int sms, double* out, long long* cyc) {
  const int iters = 20000;
  for (int warps : {4, 8, 12, 16}) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    probe<MODE><<<sms, warps * 32>>>(out, iters, 1.0, cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    probe<MODE><<<sms, warps * 32>>>(out, iters, 1.0, cyc);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    const double flops = (double)sms * warps * iters * 9 * 512.0;
    printf("mode %d warps/SM %2d (%d per sub-partition): %.2f TFLOP/s, %.1f cycles per DMMA per warp\n", MODE, warps, warps / 4, flops / ms * 1e-9,
           (double)h / (iters * 9.0));
  }
  return 0;
}
int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  double* out; long long* cyc;
  CK(cudaMalloc(&out, (size_t)prop.multiProcessorCount * 1024 * 8)); CK(cudaMalloc(&cyc, 8));
  if (run<0>(prop.multiProcessorCount, out, cyc)) return 1;
  if (run<1>(prop.multiProcessorCount, out, cyc)) return 1;
  if (run<2>(prop.multiProcessorCount, out, cyc)) return 1;
  return 0;
}
