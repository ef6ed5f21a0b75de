// Hardware micro-benchmarks that decide the FP64 stack-kernel design on B200 (sm_100a).
// Not product code: evidence for DESIGN.md.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
//   (1) DFMA chip throughput          (2) DMMA.8x8x4 chip throughput / latency
//   (3) LDS.64 wavefront cost for fragment patterns (ld=23 raw block layout)
//   (4) cp.async.bulk (UBLKCP) throughput for 4240-byte block copies, L2- and HBM-resident
//   (5) RED.ADD.F64 throughput for 529-element block flushes
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

// ---------------------------------------------------------------- (1) DFMA
template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double x, double y) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------------------------------------------------------- (2) DMMA
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void dmma_kernel(double* out, int iters, double x, double y) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c0[i] = 0; c1[i] = 0; }
  double a = x + threadIdx.x, b = y;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------------------------------------------------------- (3) LDS patterns
// pattern 0: all lanes distinct consecutive doubles; 1: DMMA fragment on ld=23 block (lane -> (lane>>2) + (lane&3)*23)
// pattern 2: fragment on ld=24; 3: 8 distinct (3*tx) broadcast over 4; 4: fragment ld=23, k strided by 4 ((lane&3)*4*23)
// pattern 5: fragment with ld=20 (conflict-free candidate)
__global__ void lds_kernel(double* out, int iters, int pattern, long long* cycles) {
  __shared__ double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  int lane = threadIdx.x & 31;
  int off;
  switch (pattern) {
    case 0: off = lane; break;
    case 1: off = (lane >> 2) + (lane & 3) * 23; break;
    case 2: off = (lane >> 2) + (lane & 3) * 24; break;
    case 3: off = 3 * (lane & 7); break;
    case 4: off = (lane >> 2) + (lane & 3) * 4 * 23; break;
    default: off = (lane >> 2) + (lane & 3) * 20; break;
  }
  double s = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      // vary the base so the compiler cannot hoist; keeps relative pattern
      s += sm[(off + u * 8 + (it & 7)) & 4095];
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// ---------------------------------------------------------------- (4) bulk copy
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "WAIT_LOOP:\n"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
    "@p bra DONE;\n"
    "bra WAIT_LOOP;\n"
    "DONE:\n"
    "}\n" ::"r"(smem_u32(bar)),
    "r"(parity)
    : "memory");
}

// Each warp owns a ring of STAGES buffers of 2 x 4240 B; lane 0 issues copies; warp waits and touches one word.
template <int STAGES>
__global__ void bulk_kernel(const char* __restrict__ src, const int* __restrict__ idx, int n_per_warp, int nblocks_total,
                            double* out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warps = blockDim.x / 32;
  const int warp = threadIdx.x / 32, lane = threadIdx.x & 31;
  const int BUF = 2 * 4352;  // 2 blocks, padded to 128B multiple
  unsigned char* base = smem_raw + (size_t)warp * STAGES * BUF;
  uint64_t* bars = (uint64_t*)(smem_raw + (size_t)warps * STAGES * BUF) + warp * STAGES;
  if (lane == 0)
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int gw = blockIdx.x * warps + warp;
  const int* my = idx + (size_t)gw * n_per_warp * 2;
  double s = 0;
  // prologue
  for (int p = 0; p < STAGES - 1 && p < n_per_warp; ++p) {
    if (lane == 0) {
      mbar_expect_tx(&bars[p], 2 * 4240);
      bulk_g2s(base + p * BUF, src + (size_t)my[2 * p] * 4240, 4240, &bars[p]);
      bulk_g2s(base + p * BUF + 4352, src + (size_t)my[2 * p + 1] * 4240, 4240, &bars[p]);
    }
  }
  for (int i = 0; i < n_per_warp; ++i) {
    const int st = i % STAGES;
    const int nx = i + STAGES - 1;
    if (nx < n_per_warp && lane == 0) {
      const int sn = nx % STAGES;
      mbar_expect_tx(&bars[sn], 2 * 4240);
      bulk_g2s(base + sn * BUF, src + (size_t)my[2 * nx] * 4240, 4240, &bars[sn]);
      bulk_g2s(base + sn * BUF + 4352, src + (size_t)my[2 * nx + 1] * 4240, 4240, &bars[sn]);
    }
    mbar_wait(&bars[st], (i / STAGES) & 1);
    s += ((double*)(base + st * BUF))[lane] + ((double*)(base + st * BUF + 4352))[lane];
    __syncwarp();
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------------------------------------------------------- (5) RED f64
__global__ void red_kernel(double* c, const int* __restrict__ blk, int n_per_warp) {
  const int warps = blockDim.x / 32;
  const int warp = threadIdx.x / 32, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * warps + warp;
  for (int i = 0; i < n_per_warp; ++i) {
    double* p = c + (size_t)blk[(size_t)gw * n_per_warp + i] * 529;
    for (int e = lane; e < 529; e += 32) atomicAdd(p + e, 1.0);
  }
}
__global__ void rmw_kernel(double* c, const int* __restrict__ blk, int n_per_warp) {
  const int warps = blockDim.x / 32;
  const int warp = threadIdx.x / 32, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * warps + warp;
  for (int i = 0; i < n_per_warp; ++i) {
    double* p = c + (size_t)blk[(size_t)gw * n_per_warp + i] * 529;
    for (int e = lane; e < 529; e += 32) p[e] += 1.0;
  }
}

template <typename F>
float time_ms(F f, int reps = 5) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a));
    f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s sms %d l2 %d MB smem/blk optin %zu clock %d kHz\n", prop.name, sms, prop.l2CacheSize >> 20,
         prop.sharedMemPerBlockOptin, prop.clockRate);
  double* out;
  CK(cudaMalloc(&out, 64 << 20));

  // (1) DFMA
  for (int warps : {4, 8, 16, 32}) {
    const int iters = 20000;
    float ms = time_ms([&] { dfma_kernel<16><<<sms, warps * 32>>>(out, iters, 1.0000001, 1e-9); });
    double flops = 2.0 * 16 * iters * (double)sms * warps * 32;
    printf("DFMA ilp16 warps/SM %2d : %.2f TFLOP/s\n", warps, flops / ms * 1e-9);
  }
  // (2) DMMA
  for (int warps : {4, 8, 16}) {
    const int iters = 4000;
    float ms = time_ms([&] { dmma_kernel<9><<<sms, warps * 32>>>(out, iters, 1.0, 1e-9); });
    double flops = 2.0 * 256 * 9 * iters * (double)sms * warps;
    printf("DMMA.8x8x4 nacc9 warps/SM %2d : %.2f TFLOP/s\n", warps, flops / ms * 1e-9);
  }
  {
    const int iters = 4000;
    float ms = time_ms([&] { dmma_kernel<1><<<sms, 4 * 32>>>(out, iters, 1.0, 1e-9); });
    printf("DMMA dependent chain: %.1f ns per DMMA (latency)\n", ms * 1e6 / iters);
    ms = time_ms([&] { dmma_kernel<2><<<sms, 4 * 32>>>(out, iters, 1.0, 1e-9); });
    printf("DMMA 2 chains: %.1f ns per pair\n", ms * 1e6 / iters);
    ms = time_ms([&] { dmma_kernel<4><<<sms, 4 * 32>>>(out, iters, 1.0, 1e-9); });
    printf("DMMA 4 chains: %.1f ns per quad\n", ms * 1e6 / iters);
  }
  // (3) LDS patterns (single warp, cycles per LDS.64)
  {
    long long* cyc;
    CK(cudaMallocManaged(&cyc, 8));
    for (int p = 0; p < 6; ++p) {
      lds_kernel<<<1, 32>>>(out, 1000, p, cyc);
      CK(cudaDeviceSynchronize());
      printf("LDS.64 pattern %d, 1 warp: %.2f cycles per load (incl. add)\n", p, (double)*cyc / (1000.0 * 16));
      lds_kernel<<<1, 256>>>(out, 1000, p, cyc);
      CK(cudaDeviceSynchronize());
      printf("LDS.64 pattern %d, 8 warps: %.2f cycles per warp-load-slot (SM pipe: /8)\n", p, (double)*cyc / (1000.0 * 16));
    }
  }
  // (4) bulk copies
  {
    const int warps = 8;
    const int STAGES = 3;
    const size_t smem = (size_t)warps * STAGES * 2 * 4352 + warps * STAGES * 8;
    CK(cudaFuncSetAttribute(bulk_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int nsrc : {10000, 200000}) {  // 42 MB (L2 resident) and 848 MB (HBM)
      char* src;
      CK(cudaMalloc(&src, (size_t)nsrc * 4240));
      CK(cudaMemset(src, 0, (size_t)nsrc * 4240));
      const int n_per_warp = 512;
      const int nw = sms * warps;
      std::vector<int> h((size_t)nw * n_per_warp * 2);
      srand(1);
      for (auto& v : h) v = rand() % nsrc;
      int* idx;
      CK(cudaMalloc(&idx, h.size() * 4));
      CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
      float ms = time_ms([&] { bulk_kernel<STAGES><<<sms, warps * 32, smem>>>(src, idx, n_per_warp, nsrc, out); });
      double bytes = (double)nw * n_per_warp * 2 * 4240;
      printf("bulk g2s 4240B x2/entry, %d src blocks (%.0f MB): %.1f GB/s, %.2f M entries/s\n", nsrc, nsrc * 4240e-6,
             bytes / ms * 1e-6, (double)nw * n_per_warp / ms * 1e-3);
      CK(cudaFree(src));
      CK(cudaFree(idx));
    }
  }
  // (5) RED / RMW flush of 529-double blocks
  {
    const int nblk = 100000;  // 423 MB
    double* c;
    CK(cudaMalloc(&c, (size_t)nblk * 529 * 8));
    CK(cudaMemset(c, 0, (size_t)nblk * 529 * 8));
    const int warps = 8, n_per_warp = 64;
    const int nw = sms * warps;
    std::vector<int> h((size_t)nw * n_per_warp);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (int)(i % nblk);
    int* blk;
    CK(cudaMalloc(&blk, h.size() * 4));
    CK(cudaMemcpy(blk, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    float ms = time_ms([&] { red_kernel<<<sms, warps * 32>>>(c, blk, n_per_warp); });
    printf("RED.f64 block flush: %.2f M blocks/s (%.1f GB/s of C)\n", h.size() / ms * 1e-3, h.size() * 4232.0 / ms * 1e-6);
    ms = time_ms([&] { rmw_kernel<<<sms, warps * 32>>>(c, blk, n_per_warp); });
    printf("LD+ST  block flush: %.2f M blocks/s (%.1f GB/s of C, x2 traffic)\n", h.size() / ms * 1e-3,
           h.size() * 4232.0 / ms * 1e-6);
  }
  printf("done\n");
  return 0;
}
