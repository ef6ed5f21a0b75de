// dbcsr_b200/csrc/smm_dmma.cuh -- FP64 stack-drain kernel for sm_100a (hand-written, B200-first).
//
// Replaces the five CUDA-core kernels of the reference (src/acc/libsmm_acc/kernels/smm_acc_dnt_{tiny,small,medium,largeDB1,
// largeDB2}.h) for every (m,n,k) with m,n <= 32.  Design (measured basis: profiles/microbench_r01.txt):
//   * the block contraction runs on the FP64 tensor pipe: DMMA.8x8x4 (mma.sync.m8n8k4.f64) reaches 37 TFLOP/s on B200 with
//     one warp per SM sub-partition, while needing 6x fewer shared-memory operand loads than a register-tiled DFMA kernel
//     (fragments are distributed one element per lane); tcgen05 has no FP64 kind, so this is the tensor path for FP64;
//   * every warp is autonomous (no CTA-wide barrier anywhere): it owns a contiguous chunk of the C-sorted stack and a private
//     ring of NST shared-memory stages; lane 0 stages the A and B blocks of entry i+NST-1 with cp.async.bulk (TMA, UBLKCP) and
//     an mbarrier transaction count while the warp multiplies entry i.  The B200 sweep (profiles/variants_r01_*.txt) settled on
//     NST = 1 with 24 resident warps per SM (small CTAs): the other warps hide the copy latency better than a deeper ring.
//     Blocks are only 8-byte aligned in the data area (4232 B for 23x23), TMA needs 16 B: the copy fetches the enclosing
//     16-byte-aligned window and the block starts `addr & 15` bytes into the stage;
//   * the raw column-major block layout is kept in shared memory (TMA cannot pad), so bank conflicts of the fragment loads
//     are removed by permuting the k index instead: DMMA sums over 4 k-values per instruction and any assignment of k to the
//     four lane groups is legal; KMap picks the stride that makes `k*ld mod 16` distinct for the four groups;
//   * C is accumulated in registers over a run of equal c_first (the stack is C-sorted) and flushed with RED.ADD.F64
//     (no read of C; L2 does the read-modify-write), which stays correct for unsorted/binned stacks and chunk boundaries;
//   * launches use programmatic dependent launch (griddepcontrol.launch_dependents at entry, .wait at exit): consecutive
//     stack drains of a stream overlap their ramp-up/tail, completion order is still stream order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace smm {

__host__ __device__ constexpr int round_up_c(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ constexpr int min_c(int a, int b) { return a < b ? a : b; }
__host__ __device__ constexpr int max_c(int a, int b) { return a > b ? a : b; }

// worst multiplicity of a 16-double-bank window when a half warp (g = 0..3, t = 0..3) loads element (k_t * ld + g),
// k_t = t * stride.  1 = conflict free (2 cycles per LDS.64), 2 = two wavefronts per half warp, ...
__host__ __device__ constexpr int frag_conflict(int ld, int stride) {
  int worst = 0;
  for (int bank = 0; bank < 16; ++bank) {
    int cnt = 0;
    for (int t = 0; t < 4; ++t)
      for (int g = 0; g < 4; ++g)
        if (((t * stride * ld + g) % 16) == bank) ++cnt;
    worst = max_c(worst, cnt);
  }
  return worst;
}

__host__ __device__ constexpr int pick_kstride(int M, int N, int K) {
  int best = 1, best_cost = 1 << 30;
  for (int s = 4; s >= 1; s /= 2) {
    if (4 * s > round_up_c(K, 4) && s > 1) continue;  // stride group does not fit into K
    const int cost = frag_conflict(M, s) + frag_conflict(N, s);
    if (cost < best_cost) {
      best_cost = cost;
      best = s;
    }
  }
  return best;
}

// k-index permutation: step s, lane group t  ->  k = kbase(s) + kstride(s) * t   (valid iff < K)
template <int K, int ST>
struct KMap {
  static constexpr int full = K / (4 * ST);
  static constexpr int rem = K - full * 4 * ST;
  static constexpr int rem_steps = (rem + 3) / 4;
  static constexpr int KS = full * ST + rem_steps;
  __host__ __device__ static constexpr int kbase(int s) { return s < full * ST ? (s / ST) * 4 * ST + (s % ST) : full * 4 * ST + (s - full * ST); }
  __host__ __device__ static constexpr int kstride(int s) { return s < full * ST ? ST : rem_steps; }
};

template <int M, int N, int K>
struct Shape {
  static constexpr int TM = (M + 7) / 8, TN = (N + 7) / 8;
  static constexpr int A_BYTES = M * K * 8, B_BYTES = N * K * 8;
  // stage buffers: 8 B possible misalignment shift + block + up to 8 doubles of (discarded) over-read by padded rows
  static constexpr int ABUF = round_up_c(8 + (M * K + 8) * 8, 128);
  static constexpr int BBUF = round_up_c(8 + (N * K + 8) * 8, 128);
  static constexpr int STAGE = ABUF + BBUF;
};

// (warps per CTA, stages).  Sweep on B200 (profiles/variants_r01_*.txt, cfg2): with programmatic dependent launch the best
// configurations use ONE stage per warp and many small CTAs per SM (8x1: 22.6, 12x1: 22.5, 4x1: 22.4 TFLOP/s) -- 24 resident
// warps hide the TMA latency better than a deeper ring under fewer warps (8x3: 19.8), and small CTAs retire/launch
// independently, which keeps the SMs busy across kernel boundaries.
__host__ __device__ constexpr int pick_nst(int stage_bytes) { return stage_bytes > 0 ? 1 : 1; }
__host__ __device__ constexpr int pick_wpc(int stage_bytes) { return (4 * stage_bytes <= 200 * 1024) ? 4 : ((2 * stage_bytes <= 200 * 1024) ? 2 : 1); }

// stack entry e = (a_first, b_first, c_first), 1-based element offsets
__device__ __forceinline__ int3 ld_entry(const int* __restrict__ stack, int e) {
  return make_int3(__ldg(stack + 3 * e), __ldg(stack + 3 * e + 1), __ldg(stack + 3 * e + 2));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
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
    "SMM_WAIT_LOOP_%=:\n"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
    "@p bra SMM_WAIT_DONE_%=;\n"
    "bra SMM_WAIT_LOOP_%=;\n"
    "SMM_WAIT_DONE_%=:\n"
    "}\n" ::"r"(smem_u32(bar)),
    "r"(parity)
    : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Stage one operand block: 16-byte-aligned window around [gaddr, gaddr+nbytes).  `limit` = end of the allocation that holds
// the data area (0 = unknown): the window may over-read up to 8 bytes past the block, which must stay inside the allocation.
// Returns the number of bytes handed to the TMA (what the mbarrier has to expect).
__device__ __forceinline__ uint32_t stage_block(unsigned char* dst, uint64_t gaddr, uint32_t nbytes, uint64_t limit, uint64_t* bar,
                                                bool issue) {
  const uint64_t src = gaddr & ~15ull;
  const uint32_t sh = (uint32_t)(gaddr & 15ull);
  uint32_t bytes = (sh + nbytes + 15u) & ~15u;
  if (limit != 0 && src + bytes > limit) {
    // last block of the allocation: copy the aligned part with the TMA, the (8-byte) tail by hand
    const uint32_t avail = (uint32_t)(limit - src) & ~15u;
    if (issue) {
      for (uint32_t o = avail; o < sh + nbytes; o += 8)
        *reinterpret_cast<double*>(dst + o) = *reinterpret_cast<const double*>(src + o);
    }
    bytes = avail;
  }
  if (issue && bytes > 0) bulk_g2s(dst, reinterpret_cast<const void*>(src), bytes, bar);
  return bytes;
}

template <int M, int N, int K, int NST, int WPC>
__global__ void __launch_bounds__(WPC * 32) smm_dmma_kernel(const int* __restrict__ stack, int stack_size, const double* __restrict__ a_data,
                                                            const double* __restrict__ b_data, double* __restrict__ c_data,
                                                            unsigned long long a_limit, unsigned long long b_limit, int chunk) {
  using SH = Shape<M, N, K>;
  constexpr int ST = pick_kstride(M, N, K);
  using KM = KMap<K, ST>;
  constexpr int TM = SH::TM, TN = SH::TN, KS = KM::KS;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int gw = blockIdx.x * WPC + warp;
  const int e0 = gw * chunk;
  const int e1 = min(e0 + chunk, stack_size);
  // Programmatic dependent launch: let the next stack kernel of this stream start filling SMs as soon as our CTAs retire
  // (stacks only accumulate into C with RED, so consecutive drains are independent); the matching wait sits at the very end so
  // that a kernel never COMPLETES before its predecessor has (later memcpys / events keep plain stream-order semantics).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (e0 >= e1) {  // warps never synchronise with each other
    asm volatile("griddepcontrol.wait;" ::: "memory");
    return;
  }

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw) + warp * NST;
  unsigned char* wbase = smem_raw + round_up_c(WPC * NST * 8, 128) + (size_t)warp * NST * SH::STAGE;

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();


  // Stack entries are fetched 32 at a time, one per lane (coalesced), and handed out with warp shuffles, so that no global
  // load sits on the per-entry critical path (ncu r01: long_scoreboard was the second largest stall). `cur` holds the entries
  // [ebase, ebase+32), `nxt` the following 32 (loaded one batch ahead).
  int ebase = e0;
  int3 cur = make_int3(1, 1, 1), nxt = make_int3(1, 1, 1);
  if (ebase + lane < e1) cur = ld_entry(stack, ebase + lane);
  if (ebase + 32 + lane < e1) nxt = ld_entry(stack, ebase + 32 + lane);
  auto entry = [&](int e) -> int3 {  // warp-uniform e in [ebase, ebase+64)
    const int r = e - ebase;
    const int3 a = make_int3(__shfl_sync(0xffffffffu, cur.x, r & 31), __shfl_sync(0xffffffffu, cur.y, r & 31),
                             __shfl_sync(0xffffffffu, cur.z, r & 31));
    const int3 b = make_int3(__shfl_sync(0xffffffffu, nxt.x, r & 31), __shfl_sync(0xffffffffu, nxt.y, r & 31),
                             __shfl_sync(0xffffffffu, nxt.z, r & 31));
    return r < 32 ? a : b;
  };

  auto issue = [&](int e) {  // executed by the whole warp (uniform), the copies are issued by lane 0
    const int3 p = entry(e);
#if defined(SMM_ABL_NOTMA)
    if (p.x < 0) {
#else
    if (lane == 0) {
#endif
      const int sidx = (e - e0) % NST;
      unsigned char* stg = wbase + (size_t)sidx * SH::STAGE;
      const uint64_t ga = reinterpret_cast<uint64_t>(a_data + (p.x - 1));
      const uint64_t gb = reinterpret_cast<uint64_t>(b_data + (p.y - 1));
      // expect first (the count must be known before the copies can complete), then copy
      const uint32_t ba = stage_block(stg, ga, SH::A_BYTES, a_limit, &bars[sidx], false);
      const uint32_t bb = stage_block(stg + SH::ABUF, gb, SH::B_BYTES, b_limit, &bars[sidx], false);
      mbar_expect_tx(&bars[sidx], ba + bb);
      stage_block(stg, ga, SH::A_BYTES, a_limit, &bars[sidx], true);
      stage_block(stg + SH::ABUF, gb, SH::B_BYTES, b_limit, &bars[sidx], true);
    }
  };

#pragma unroll
  for (int p = 0; p < NST - 1; ++p)
    if (e0 + p < e1) issue(e0 + p);

  double acc[TM][TN][2];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto flush = [&](int c_first) {
#if defined(SMM_ABL_NOFLUSH)
    if (c_first > 0) {  // ablation: no RED traffic (results wrong)
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j)
          if (acc[i][j][0] == 1.2345e300) c_data[0] = acc[i][j][1];
      return;
    }
#endif
    double* __restrict__ cb = c_data + (c_first - 1);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int row = i * 8 + g;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int col = j * 8 + 2 * t;
        if ((i * 8 + 7 < M) || (row < M)) {
          if ((j * 8 + 7 < N) || (col < N)) atomicAdd(cb + col * M + row, acc[i][j][0]);
          if ((j * 8 + 7 < N) || (col + 1 < N)) atomicAdd(cb + (col + 1) * M + row, acc[i][j][1]);
        }
        acc[i][j][0] = acc[i][j][1] = 0.0;
      }
    }
  };

  int cur_c = -1;
  for (int e = e0; e < e1; ++e) {
    if (e - ebase >= 32) {  // roll the entry batches; the new `nxt` is needed 32 - (NST-1) entries from now at the earliest
      ebase += 32;
      cur = nxt;
      nxt = make_int3(1, 1, 1);
      if (ebase + 32 + lane < e1) nxt = ld_entry(stack, ebase + 32 + lane);
    }
    // stage (e-1)%NST was consumed by the previous iteration (guarded by the __syncwarp at its end): refill it
    if (e + NST - 1 < e1) issue(e + NST - 1);

    const int3 p = entry(e);
    if (p.z != cur_c) {
      if (cur_c >= 0) flush(cur_c);
      cur_c = p.z;
    }
    const int i = e - e0;
    const int sidx = i % NST;
    const unsigned char* stg = wbase + (size_t)sidx * SH::STAGE;
    const uint32_t sha = (uint32_t)(reinterpret_cast<uint64_t>(a_data + (p.x - 1)) & 15ull);
    const uint32_t shb = (uint32_t)(reinterpret_cast<uint64_t>(b_data + (p.y - 1)) & 15ull);
    const double* __restrict__ As = reinterpret_cast<const double*>(stg + sha);
    const double* __restrict__ Bs = reinterpret_cast<const double*>(stg + SH::ABUF + shb);

#if !defined(SMM_ABL_NOTMA)
    mbar_wait(&bars[sidx], (uint32_t)((i / NST) & 1));
#endif

#pragma unroll
    for (int s = 0; s < KS; ++s) {
      const int kb = KM::kbase(s), kst = KM::kstride(s);
      const int k = kb + kst * t;
      const bool all_valid = (kb + 3 * kst < K);
      const bool valid = all_valid || (k < K);
      double af[TM], bf[TN];
#pragma unroll
#if defined(SMM_ABL_NOLDS)
      for (int ti = 0; ti < TM; ++ti) af[ti] = (double)(s + ti + sha);
#pragma unroll
      for (int tj = 0; tj < TN; ++tj) bf[tj] = (double)(s - tj + shb);
      (void)As; (void)Bs; (void)valid;
#else
      for (int ti = 0; ti < TM; ++ti) af[ti] = valid ? As[k * M + ti * 8 + g] : 0.0;
#pragma unroll
      for (int tj = 0; tj < TN; ++tj) bf[tj] = valid ? Bs[k * N + tj * 8 + g] : 0.0;
#endif
#pragma unroll
      for (int ti = 0; ti < TM; ++ti)
#pragma unroll
        for (int tj = 0; tj < TN; ++tj) dmma884(acc[ti][tj][0], acc[ti][tj][1], af[ti], bf[tj]);
    }
    __syncwarp();  // every lane is done reading this stage before lane 0 refills it
  }
  if (cur_c >= 0) flush(cur_c);
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

}  // namespace smm
