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
//   * C is accumulated in registers over a run of equal c_first (the stack is C-sorted) and flushed by an L2-side reduction
//     (no read of C), which stays correct for unsorted/binned stacks, chunk boundaries and overlapping launches: either one
//     RED.ADD.F64 per element, or -- FLUSH 2, shipped for 23^3 -- an image of the run in the warp's free operand stage that ONE
//     cp.reduce.async.bulk.add.f64 (TMA, UBLKRED) adds into C.  Runs are short on real stacks (1.7 entries on the 23^3/10 %
//     workload), so the flush is on the critical resource list: per-element REDs leave an SM at ~1 element per cycle;
//   * launches use programmatic dependent launch: griddepcontrol.launch_dependents at entry lets the next kernel of the stream
//     become resident early.  By default the kernel then executes griddepcontrol.wait BEFORE its first global read (the
//     predecessor may be a producer of A/B/C/stack: memset, transpose, pack -- only the launch latency and the prologue
//     overlap).  When the caller has declared the stream a chain of independent stack drains (FLAG_PDL_CHAIN, set through
//     libsmm_acc_b200_stream_chain), the wait moves to the very end: consecutive drains overlap their ramp-up/tail and
//     completion order is still stream order.
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
// L2 cache-policy variants (HINT template parameter of the kernels): operand blocks are re-used ~100x by other warps within a few
// consecutive stacks and should outlive the C lines, which one launch touches once or twice (evict_first on the RED).
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                 smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void red_add_hint(double* addr, double v, uint64_t pol) {
  asm volatile("red.global.add.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(addr), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ unsigned long long clock_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
  return t;
}
__device__ __forceinline__ unsigned long long globaltimer_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
  return t;
}
__device__ __forceinline__ unsigned smid_now() {
  unsigned r;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
  return r;
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {  // non-blocking
  uint32_t ok;
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
    "selp.u32 %0, 1, 0, p;\n"
    "}\n"
    : "=r"(ok)
    : "r"(smem_u32(bar)), "r"(parity)
    : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
template <int HINT>
__device__ __forceinline__ uint32_t stage_block(unsigned char* dst, uint64_t gaddr, uint32_t nbytes, uint64_t limit, uint64_t* bar,
                                                bool issue, uint64_t pol) {
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
  if (issue && bytes > 0) {
    if (HINT >= 2)
      bulk_g2s_hint(dst, reinterpret_cast<const void*>(src), bytes, bar, pol);
    else
      bulk_g2s(dst, reinterpret_cast<const void*>(src), bytes, bar);
  }
  return bytes;
}

// Ablation bits (ABL template parameter; experiment builds only -- results are WRONG with any bit set)
constexpr int ABL_NOFLUSH = 1, ABL_NOLDS = 2, ABL_NOTMA = 4;

// One stack entry on the FP64 tensor pipe: acc += A(M x K, col-major, ld M) * Bt(N x K, col-major, ld N)^T, both in shared
// memory; g = lane / 4 (row of the fragment), t = lane % 4 (k of the fragment); k permuted by KMap (see the file header).
template <int M, int N, int K, int ABL = 0>
__device__ __forceinline__ void mma_entry(const double* __restrict__ As, const double* __restrict__ Bs,
                                          double (&acc)[Shape<M, N, K>::TM][Shape<M, N, K>::TN][2], int g, int t, uint32_t sha,
                                          uint32_t shb) {
  using SH = Shape<M, N, K>;
  constexpr int ST = pick_kstride(M, N, K);
  using KM = KMap<K, ST>;
  constexpr int TM = SH::TM, TN = SH::TN, KS = KM::KS;
#pragma unroll
  for (int s = 0; s < KS; ++s) {
    const int kb = KM::kbase(s), kst = KM::kstride(s);
    const int k = kb + kst * t;
    const bool all_valid = (kb + 3 * kst < K);
    const bool valid = all_valid || (k < K);
    double af[TM], bf[TN];
    if constexpr ((ABL & ABL_NOLDS) != 0) {
#pragma unroll
      for (int ti = 0; ti < TM; ++ti) af[ti] = (double)(s + ti + sha);
#pragma unroll
      for (int tj = 0; tj < TN; ++tj) bf[tj] = (double)(s - tj + shb);
    }
    else {
#pragma unroll
      for (int ti = 0; ti < TM; ++ti) af[ti] = valid ? As[k * M + ti * 8 + g] : 0.0;
#pragma unroll
      for (int tj = 0; tj < TN; ++tj) bf[tj] = valid ? Bs[k * N + tj * 8 + g] : 0.0;
    }
#pragma unroll
    for (int ti = 0; ti < TM; ++ti)
#pragma unroll
      for (int tj = 0; tj < TN; ++tj) dmma884(acc[ti][tj][0], acc[ti][tj][1], af[ti], bf[tj]);
  }
}

// Flush a run's accumulators into the C block at 1-based element offset c_first with RED.ADD.F64 and clear them.
template <int M, int N, int K, int HINT, int ABL = 0>
__device__ __forceinline__ void flush_acc(double* __restrict__ c_data, int c_first, double (&acc)[Shape<M, N, K>::TM][Shape<M, N, K>::TN][2],
                                          int g, int t, uint64_t pol_c) {
  constexpr int TM = Shape<M, N, K>::TM, TN = Shape<M, N, K>::TN;
  if constexpr ((ABL & ABL_NOFLUSH) != 0) {
    if (c_first > 0) {  // ablation: no RED traffic (results wrong)
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j)
          if (acc[i][j][0] == 1.2345e300) c_data[0] = acc[i][j][1];
      return;
    }
  }
  double* __restrict__ cb = c_data + (c_first - 1);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = i * 8 + g;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int col = j * 8 + 2 * t;
      if ((i * 8 + 7 < M) || (row < M)) {
        if ((j * 8 + 7 < N) || (col < N)) {
          if (HINT >= 1)
            red_add_hint(cb + col * M + row, acc[i][j][0], pol_c);
          else
            atomicAdd(cb + col * M + row, acc[i][j][0]);
        }
        if ((j * 8 + 7 < N) || (col + 1 < N)) {
          if (HINT >= 1)
            red_add_hint(cb + (col + 1) * M + row, acc[i][j][1], pol_c);
          else
            atomicAdd(cb + (col + 1) * M + row, acc[i][j][1]);
        }
      }
      acc[i][j][0] = acc[i][j][1] = 0.0;
    }
  }
}

// Flush through the TMA: the accumulators are written to a per-warp shared-memory image of the C block (same 16-byte phase
// as the block has in global memory) and ONE cp.reduce.async.bulk (.add.f64, SASS UBLKRED) adds the 16-byte-aligned middle
// of the block into C inside L2; the at most two 8-byte end pieces of an odd-aligned block go out as scalar REDs.  This
// takes the M*N per-element REDs of flush_acc off the SM's load/store path (about one element per cycle per SM).
// The caller must have waited for the previous bulk reduction of this scratch buffer (cp.async.bulk.wait_group.read).
__host__ __device__ constexpr int scratch_bytes(int M, int N) { return round_up_c(16 + M * N * 8, 128); }

template <int M, int N, int K>
__device__ __forceinline__ void flush_acc_bulk(double* __restrict__ c_data, int c_first, double (&acc)[Shape<M, N, K>::TM][Shape<M, N, K>::TN][2],
                                               int g, int t, int lane, unsigned char* scratch) {
  constexpr int TM = Shape<M, N, K>::TM, TN = Shape<M, N, K>::TN;
  const uint64_t gc = reinterpret_cast<uint64_t>(c_data + (c_first - 1));
  const uint32_t sh = (uint32_t)(gc & 15ull);
  double* __restrict__ sc = reinterpret_cast<double*>(scratch + sh);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = i * 8 + g;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int col = j * 8 + 2 * t;
      if ((i * 8 + 7 < M) || (row < M)) {
        if ((j * 8 + 7 < N) || (col < N)) sc[col * M + row] = acc[i][j][0];
        if ((j * 8 + 7 < N) || (col + 1 < N)) sc[(col + 1) * M + row] = acc[i][j][1];
      }
      acc[i][j][0] = acc[i][j][1] = 0.0;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk (async proxy) read
  __syncwarp();
  if (lane == 0) {
    const uint32_t head = (16u - sh) & 15u;            // bytes in front of the first 16-byte boundary (0 or 8)
    const uint32_t end = sh + (uint32_t)(M * N * 8);   // block end relative to the scratch base
    const uint32_t mid0 = sh + head, mid1 = end & ~15u;
    if (mid1 > mid0) {
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(gc + head),
                   "r"(smem_u32(scratch + mid0)), "r"(mid1 - mid0)
                   : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (head != 0) atomicAdd(reinterpret_cast<double*>(gc), sc[0]);
    if ((end & 15u) != 0 && (M * N > 1 || head == 0)) atomicAdd(reinterpret_cast<double*>(gc) + (M * N - 1), sc[M * N - 1]);
  }
}

// Chunk of global warp gw: legacy split (extra < 0: `chunk` entries per warp, the last warps idle) or balanced split
// (extra >= 0: `chunk` = floor(S / warps) entries, the first `extra` = S mod warps warps take one more).
__device__ __forceinline__ void warp_chunk(int gw, int chunk, int extra, int stack_size, int& e0, int& e1) {
  if (extra < 0) {
    e0 = min(gw * chunk, stack_size);
    e1 = min(e0 + chunk, stack_size);
  }
  else {
    e0 = min(gw * chunk + min(gw, extra), stack_size);
    e1 = min(e0 + chunk + (gw < extra ? 1 : 0), stack_size);
  }
}

// Kernel flags
constexpr int FLAG_ALIGN_RUNS = 1;  // move chunk boundaries to the next change of c_first (at most ALIGN_LOOKAHEAD - 1 entries ahead)
// How far a chunk boundary may move.  Round 1 looked 30 entries ahead: fine for DBCSR's stacks (mean run 1.7 entries), but on
// stacks with LONG runs (the device builder's tile order: all ~10 products of a C block adjacent) boundaries snapped from 12-entry
// chunks to alternating 10 / 20: the slowest warp of a CTA decides when its slot is free again (9.87 ms per multiply against 8.17
// with alignment off).  With 4, short runs are still kept whole and a long run is simply split (one extra flush).
constexpr int ALIGN_LOOKAHEAD = 4;
constexpr int FLAG_PDL_CHAIN = 2;   // the predecessor in the stream is an independent stack drain: no grid dependency before the reads

// Trace record of one warp (TRACE kernels; lane 0 writes): see tools/trace_analyze.py
//   [0] smid | n_entries << 32   [1] globaltimer at start   [2] clock at start   [3] first entry
//   [4 + 4 i + {0,1,2,3}], i < 30: clock after the copies of entry i were issued (warp-specialised kernel: clock before the
//                                  wait), clock when its operands had arrived, clock after its last DMMA was issued, c_first
//   [124] number of flushes  [125] cycles spent issuing flushes  [126] clock at end  [127] globaltimer at end
constexpr int TRACE_WORDS = 128;
constexpr int TRACE_ENTRIES = 30;

// FLUSH: 0 = per-element RED (flush_acc), 1 = bulk reduction through the TMA (flush_acc_bulk) from a scratch buffer per warp,
//        2 = bulk reduction from the warp's (single) operand stage, which is free between two entries: no extra shared memory,
//            but the copies of the next entry are issued only after the bulk engine has read the stage
//        4 = one scratch image per CTA, taken under a shared-memory lock: the copies of the next entry are issued before the
//            flush as in the RED kernel and the price is one image instead of one per warp (experiment variant, not yet measured)
//        3 = like 2 with the image confined to the A half of the stage, so that the B copy of the next entry is issued BEFORE
//            the flush and only the A copy waits for the bulk engine (experiment variant, not yet measured)
template <int M, int N, int K, int NST, int WPC, int FLUSH>
struct BaseGeom {
  using SH = Shape<M, N, K>;
  static constexpr int BAR_BYTES = round_up_c(WPC * NST * 8, 128);
  static constexpr int SCRATCH = FLUSH == 1 ? scratch_bytes(M, N) : 0;
  static constexpr int PER_WARP = NST * SH::STAGE + SCRATCH;
  // FLUSH 4: ONE scratch image per CTA (behind the warps' stages) + a lock word in front of it
  static constexpr int CTA_SCRATCH = FLUSH == 4 ? 128 + scratch_bytes(M, N) : 0;
  static constexpr int SMEM = BAR_BYTES + WPC * PER_WARP + CTA_SCRATCH;
};

template <int M, int N, int K, int NST, int WPC, int HINT = 0, bool TRACE = false, int FLUSH = 0, int ABL = 0>
__global__ void __launch_bounds__(WPC * 32) smm_dmma_kernel(const int* __restrict__ stack, int stack_size, const double* __restrict__ a_data,
                                                            const double* __restrict__ b_data, double* __restrict__ c_data,
                                                            unsigned long long a_limit, unsigned long long b_limit, int chunk, int extra,
                                                            int flags, unsigned long long* __restrict__ trace) {
  using SH = Shape<M, N, K>;
  using G = BaseGeom<M, N, K, NST, WPC, FLUSH>;
  constexpr int TM = SH::TM, TN = SH::TN;
  static_assert(FLUSH != 2 || (NST == 1 && SH::STAGE >= scratch_bytes(M, N)), "FLUSH 2 stages the C image in the single operand stage");
  static_assert(FLUSH != 3 || (NST == 1 && SH::ABUF >= scratch_bytes(M, N)), "FLUSH 3 stages the C image in the A half of the stage");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int gw = blockIdx.x * WPC + warp;
  int n0, n1;  // nominal chunk
  warp_chunk(gw, chunk, extra, stack_size, n0, n1);
  // Programmatic dependent launch: let the next stack kernel of this stream start filling SMs as soon as our CTAs retire.
  // Unless the caller declared the stream a chain of independent drains (stacks only accumulate into C with RED), everything
  // the predecessor wrote must be complete and visible before the first global read below.  In chain mode the matching wait
  // sits at the very end so that a kernel never COMPLETES before its predecessor has (later memcpys / events keep plain
  // stream-order semantics).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if ((flags & FLAG_PDL_CHAIN) == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
  if constexpr (FLUSH == 4) {  // the only CTA-wide barrier of this kernel: the scratch lock starts open (before any warp leaves)
    if (threadIdx.x == 0) *reinterpret_cast<int*>(smem_raw + G::BAR_BYTES + (size_t)WPC * G::PER_WARP) = 0;
    __syncthreads();
  }
  if (n0 >= n1) {  // warps never synchronise with each other
    asm volatile("griddepcontrol.wait;" ::: "memory");
    return;
  }

  // Stack entries are fetched 32 at a time, one per lane (coalesced), and handed out with warp shuffles, so that no global
  // load sits on the per-entry critical path (ncu r01: long_scoreboard was the second largest stall). `cur` holds the entries
  // [ebase, ebase+32), `nxt` the following 32 (loaded one batch ahead).
  int ebase = n0;
  int3 cur = make_int3(1, 1, 1), nxt = make_int3(1, 1, 1);
  if (ebase + lane < stack_size) cur = ld_entry(stack, ebase + lane);
  if (ebase + 32 + lane < stack_size) nxt = ld_entry(stack, ebase + 32 + lane);

  // Run-aligned chunk boundaries: boundary(n) = first j in [n, n+30] with c_first[j] != c_first[n-1] (j = S counts), else n.
  // Both neighbours evaluate the same function, so the chunks still tile the stack; a run that would be split between two
  // warps (two flushes) is drained by one.  Correctness never depends on it: C is accumulated with RED either way.
  int e0 = n0, e1 = n1;
  if ((flags & FLAG_ALIGN_RUNS) != 0) {
    int c_prev = 0, w1 = -1;
    if (n0 > 0) c_prev = __ldg(stack + 3 * (n0 - 1) + 2);
    if (n1 < stack_size && n1 - 1 + lane < stack_size) w1 = __ldg(stack + 3 * (n1 - 1 + lane) + 2);
    if (n0 > 0) {
      const bool differs = (lane < ALIGN_LOOKAHEAD) && ((n0 + lane >= stack_size) || (cur.z != c_prev));
      const unsigned m = __ballot_sync(0xffffffffu, differs);
      if (m != 0) e0 = n0 + (__ffs(m) - 1);
    }
    if (n1 < stack_size) {
      const int w0 = __shfl_sync(0xffffffffu, w1, 0);
      const bool differs = (lane >= 1) && (lane <= ALIGN_LOOKAHEAD) && ((n1 - 1 + lane >= stack_size) || (w1 != w0));
      const unsigned m = __ballot_sync(0xffffffffu, differs);
      if (m != 0) e1 = n1 - 1 + (__ffs(m) - 1);
    }
    if (e0 >= e1) {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      return;
    }
  }

  unsigned long long* rec = nullptr;
  unsigned long long t_flush = 0, n_flush = 0;
  if (TRACE) {
    if (trace != nullptr && gw < 4096 && lane == 0) {
      rec = trace + (size_t)gw * TRACE_WORDS;
      rec[0] = (unsigned long long)smid_now() | ((unsigned long long)(e1 - e0) << 32);
      rec[1] = globaltimer_now();
      rec[2] = clock_now();
      rec[3] = (unsigned long long)e0;
    }
  }

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw) + warp * NST;
  unsigned char* wbase = smem_raw + G::BAR_BYTES + (size_t)warp * G::PER_WARP;
  unsigned char* scratch = wbase + NST * SH::STAGE;  // FLUSH == 1 only

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint64_t pol_ab = 0, pol_c = 0;
  if (HINT >= 1) pol_c = policy_evict_first();
  if (HINT >= 2) pol_ab = policy_evict_last();

  auto entry = [&](int e) -> int3 {  // warp-uniform e in [ebase, ebase+64)
    const int r = e - ebase;
    const int3 a = make_int3(__shfl_sync(0xffffffffu, cur.x, r & 31), __shfl_sync(0xffffffffu, cur.y, r & 31),
                             __shfl_sync(0xffffffffu, cur.z, r & 31));
    const int3 b = make_int3(__shfl_sync(0xffffffffu, nxt.x, r & 31), __shfl_sync(0xffffffffu, nxt.y, r & 31),
                             __shfl_sync(0xffffffffu, nxt.z, r & 31));
    return r < 32 ? a : b;
  };

  // parts: 1 = arm the barrier with the byte count of both blocks + copy the B block, 2 = copy the A block, 3 = both
  auto issue_parts = [&](int e, int parts) {  // executed by the whole warp (uniform), the copies are issued by lane 0
    const int3 p = entry(e);
    if (((ABL & ABL_NOTMA) != 0) ? (p.x < 0) : (lane == 0)) {
      const int sidx = (e - e0) % NST;
      unsigned char* stg = wbase + (size_t)sidx * SH::STAGE;
      const uint64_t ga = reinterpret_cast<uint64_t>(a_data + (p.x - 1));
      const uint64_t gb = reinterpret_cast<uint64_t>(b_data + (p.y - 1));
      if ((parts & 1) != 0) {
        // expect first (the count must be known before the copies can complete), then copy
        const uint32_t ba = stage_block<HINT>(stg, ga, SH::A_BYTES, a_limit, &bars[sidx], false, pol_ab);
        const uint32_t bb = stage_block<HINT>(stg + SH::ABUF, gb, SH::B_BYTES, b_limit, &bars[sidx], false, pol_ab);
        mbar_expect_tx(&bars[sidx], ba + bb);
        stage_block<HINT>(stg + SH::ABUF, gb, SH::B_BYTES, b_limit, &bars[sidx], true, pol_ab);
      }
      if ((parts & 2) != 0) {
        stage_block<HINT>(stg, ga, SH::A_BYTES, a_limit, &bars[sidx], true, pol_ab);
        if (TRACE) {
          if (rec != nullptr && e - e0 < TRACE_ENTRIES) rec[4 + 4 * (e - e0)] = clock_now();
        }
      }
    }
  };
  auto issue = [&](int e) { issue_parts(e, 3); };

#pragma unroll
  for (int p = 0; p < NST - 1; ++p)
    if (e0 + p < e1) issue(e0 + p);

  double acc[TM][TN][2];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto flush = [&](int c_first) {
    unsigned long long tf0 = 0;
    if (TRACE) tf0 = clock_now();
    if constexpr (FLUSH == 1) {
      // the previous bulk reduction must have finished READING the scratch image before it is overwritten
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      flush_acc_bulk<M, N, K>(c_data, c_first, acc, g, t, lane, scratch);
    }
    else if constexpr (FLUSH == 4) {
      // CTA-wide scratch image under a spin lock (the holder never waits for another warp, so this cannot deadlock)
      unsigned char* cta_scratch = smem_raw + G::BAR_BYTES + (size_t)WPC * G::PER_WARP;
      int* lock = reinterpret_cast<int*>(cta_scratch);
      if (lane == 0) {
        while (atomicCAS(lock, 0, 1) != 0) __nanosleep(64);
        __threadfence_block();
      }
      __syncwarp();
      flush_acc_bulk<M, N, K>(c_data, c_first, acc, g, t, lane, cta_scratch + 128);
      if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the image has been read: the next holder may overwrite it
        __threadfence_block();
        atomicExch(lock, 0);
      }
      __syncwarp();
    }
    else if constexpr (FLUSH == 2 || FLUSH == 3) {
      flush_acc_bulk<M, N, K>(c_data, c_first, acc, g, t, lane, wbase);
      // lane 0 issues the next entry's copies into this stage: only after the bulk engine has read the C image
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    else {
      flush_acc<M, N, K, HINT, ABL>(c_data, c_first, acc, g, t, pol_c);
    }
    if (TRACE) {
      t_flush += clock_now() - tf0;
      ++n_flush;
    }
  };

  int cur_c = -1;
  for (int e = e0; e < e1; ++e) {
    if (e - ebase >= 32) {  // roll the entry batches; the new `nxt` is needed 32 - (NST-1) entries from now at the earliest
      ebase += 32;
      cur = nxt;
      nxt = make_int3(1, 1, 1);
      if (ebase + 32 + lane < stack_size) nxt = ld_entry(stack, ebase + 32 + lane);
    }
    // stage (e-1)%NST was consumed by the previous iteration (guarded by the __syncwarp at its end): refill it
    if constexpr (FLUSH != 2 && FLUSH != 3) {
      if (e + NST - 1 < e1) issue(e + NST - 1);
    }
    const int3 p = entry(e);
    if constexpr (FLUSH == 3) {
      // the C image only occupies the A half of the stage: the B block of this entry is already on its way while the bulk
      // engine reads the image; only the A copy has to wait for it
      if (p.z != cur_c && cur_c >= 0) {
        issue_parts(e, 1);
        flush(cur_c);
        issue_parts(e, 2);
      }
      else {
        issue(e);
      }
      cur_c = p.z;
    }
    else {
      if (p.z != cur_c) {
        if (cur_c >= 0) flush(cur_c);
        cur_c = p.z;
      }
      if constexpr (FLUSH == 2) issue(e);  // after the flush, which borrows the stage
    }
    const int i = e - e0;
    const int sidx = i % NST;
    const unsigned char* stg = wbase + (size_t)sidx * SH::STAGE;
    const uint32_t sha = (uint32_t)(reinterpret_cast<uint64_t>(a_data + (p.x - 1)) & 15ull);
    const uint32_t shb = (uint32_t)(reinterpret_cast<uint64_t>(b_data + (p.y - 1)) & 15ull);
    const double* __restrict__ As = reinterpret_cast<const double*>(stg + sha);
    const double* __restrict__ Bs = reinterpret_cast<const double*>(stg + SH::ABUF + shb);

    if constexpr ((ABL & ABL_NOTMA) == 0) mbar_wait(&bars[sidx], (uint32_t)((i / NST) & 1));
    if (TRACE) {
      if (rec != nullptr && i < TRACE_ENTRIES) {
        rec[4 + 4 * i + 1] = clock_now();
        rec[4 + 4 * i + 3] = (unsigned long long)p.z;
      }
    }

    mma_entry<M, N, K, ABL>(As, Bs, acc, g, t, sha, shb);
    __syncwarp();  // every lane is done reading this stage before lane 0 refills it
    if (TRACE) {
      if (rec != nullptr && i < TRACE_ENTRIES) rec[4 + 4 * i + 2] = clock_now();
    }
  }
  if (cur_c >= 0) flush(cur_c);
  if constexpr (FLUSH != 0) {
    // all bulk reductions of this warp are performed before it exits (shared memory is released, C must be complete)
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  if (TRACE) {
    if (rec != nullptr) {
      rec[124] = n_flush;
      rec[125] = t_flush;
      rec[126] = clock_now();
      rec[127] = globaltimer_now();
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

}  // namespace smm
