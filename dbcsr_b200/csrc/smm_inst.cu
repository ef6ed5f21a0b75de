// dbcsr_b200/csrc/smm_inst.cu -- instantiates smm_dmma_kernel<SMM_M, n, k> for every (n,k) of the tuned size set and
// exports smm::lookup_m<SMM_M>(n,k).  Compiled once per SMM_M (see Makefile) so the 125 kernels build in parallel.
// This is the ahead-of-time replacement of the reference's NVRTC JIT + kernel cache (src/acc/libsmm_acc/libsmm_acc.cpp:90-253).
#include <atomic>
#include <cstdio>
#include <cstdlib>

#include "smm_dmma.cuh"
#if defined(SMM_EXPERIMENT)
#  include "smm_dmma_ws.cuh"  // warp-specialised variant: a measured-slower experiment, never part of the shipped library
#endif
#include "smm_launch.h"
#include "smm_tiny.cuh"
#include "smm_tune.h"

#ifndef SMM_M
#  error "compile with -DSMM_M=<block rows>"
#endif

namespace smm {
// Per-shape launch policy of the warp-autonomous kernel, from the B200 sweeps with tools/kbench (profiles/r01_kbench_sweep*.txt,
// cfg2-like stacks of 30000 entries, TFLOP/s kernel-only):
//   23x23x23: RED flush, one wave of 9-entry chunks 22.6 | run-aligned chunks 23.1 | bulk flush from a scratch buffer (16 warps
//             per SM) 24.5 | bulk flush from the operand stage (24 warps) 25.2 | + run-aligned 25.6 | + 12-entry chunks 26.2
//   32^3, 26^3, 13^3, 5^3: the bulk flush and the aligned chunks are neutral or slower (few resident warps / tiny blocks), so
//             they keep the RED flush and one wave of equal chunks.
// FLUSH: see smm_dmma.cuh; CHUNK: entries per warp when the stack is large enough to give every SM a CTA (0 = one resident
// wave); ALIGN: chunk boundaries moved to changes of c_first.  The run-time knobs (smm_tune.h) override CHUNK / ALIGN when >= 0.
template <int M, int N, int K>
struct Policy {
  static constexpr int ALGO = 0;  // 0 = warp-autonomous DMMA + TMA kernel (smm_dmma.cuh), 1 = lane-per-element kernel (smm_tiny.cuh)
  static constexpr int WPC = 0;   // warps per CTA, 0 = the kernel's default (pick_wpc / 8 for the tiny kernel)
  static constexpr int FLUSH = 0, CHUNK = 0;
  static constexpr bool ALIGN = false;
};
// the tuned records: generated from the autotune database dbcsr_b200/parameters/parameters_B200.json (tools/gen_policy.py),
// like the reference generates parameters.h from parameters_<GPU>.json
#define SMM_POLICY(M_, N_, K_, ALGO_, WPC_, FLUSH_, CHUNK_, ALIGN_) \
  template <>                                                        \
  struct Policy<M_, N_, K_> {                                        \
    static constexpr int ALGO = ALGO_, WPC = WPC_;                   \
    static constexpr int FLUSH = FLUSH_, CHUNK = CHUNK_;             \
    static constexpr bool ALIGN = ALIGN_;                            \
  };
#include "smm_policy.inc"
#undef SMM_POLICY

namespace {

// DBCSR_B200_PDL=0 disables programmatic dependent launch (A/B experiments); default on
const bool g_use_pdl = [] {
  const char* e = getenv("DBCSR_B200_PDL");
  return e == nullptr || atoi(e) != 0;
}();

// trace slot of this launch (TRACE kernels only): nullptr outside the recording window set by libsmm_acc_b200_set_tunable
unsigned long long* trace_slot() {
  const int seq = g_tune.seq.fetch_add(1, std::memory_order_relaxed);
  unsigned long long* t = g_tune.trace.load(std::memory_order_relaxed);
  if (t == nullptr) return nullptr;
  const int first = g_tune.trace_first.load(std::memory_order_relaxed), cnt = g_tune.trace_count.load(std::memory_order_relaxed);
  if (seq < first || seq >= first + cnt) return nullptr;
  return t + (size_t)(seq - first) * TRACE_WARPS * TRACE_REC;
}

// cudaFuncSetAttribute and the occupancy are PER DEVICE (DBCSR picks the device per rank, c_dbcsr_acc_set_active_device may switch
// it): caches are arrays indexed by the active device, 0 = not yet set up on that device.
constexpr int kMaxDevices = 64;
struct DevCache {
  std::atomic<int> ctas_per_sm[kMaxDevices];
  std::atomic<int> num_sms[kMaxDevices];
};

template <typename Kern>
int occupancy(Kern kern, int threads, int smem, DevCache& cache, int& num_sms) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -30;
  int cps = cache.ctas_per_sm[dev].load(std::memory_order_acquire);
  if (cps == 0) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -30;
    int nb = 0, sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) return -30;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem) != cudaSuccess || nb < 1) return -30;
    cache.num_sms[dev].store(sms, std::memory_order_relaxed);
    cps = nb;
    cache.ctas_per_sm[dev].store(cps, std::memory_order_release);
  }
  num_sms = cache.num_sms[dev].load(std::memory_order_relaxed);
  return cps;
}

template <typename Kern, typename... Args>
int launch_pdl(Kern kern, int grid, int threads, int smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t err = cudaLaunchKernelEx(&cfg, kern, args...);
  return (err == cudaSuccess) ? 0 : -31;
}

// warp-autonomous kernel (smm_dmma.cuh): NST stages per warp, WPC warps per CTA
template <int M, int N, int K, int NST, int WPC, int HINT, bool TRACE, int FLUSH = 0, int ABL = 0>
int launch_base(const int* dev_stack, int stack_size, const double* a, const double* b, double* c, uint64_t a_limit, uint64_t b_limit,
                cudaStream_t stream) {
  using G = BaseGeom<M, N, K, NST, WPC, FLUSH>;
  constexpr int SMEM = G::SMEM;
  static_assert(SMEM <= 227 * 1024, "shared memory budget exceeded");
  auto kern = smm_dmma_kernel<M, N, K, NST, WPC, HINT, TRACE, FLUSH, ABL>;
  static DevCache cache;  // zero-initialised
  int g_num_sms = 0;
  const int cps = occupancy(kern, WPC * 32, SMEM, cache, g_num_sms);
  if (cps < 0) return cps;
  if (stack_size <= 0) return 0;  // stack_size 0 = "prepare": the attribute is set, nothing is launched
  const int max_grid = g_num_sms * cps;
  // at least 4 entries per warp so that the pipeline prologue and the C flush are amortised
  int grid = (stack_size + WPC * 4 - 1) / (WPC * 4);
  // chain mode (libsmm_acc_b200_stream_chain): the previous drain of the stream is still running when this one starts
  const bool chained = g_use_pdl && stream_chain_mode(stream);
  int max_chunk = g_tune.chunk.load(std::memory_order_relaxed);
  if (max_chunk < 0) {
    // per-shape policy: CHUNK entries per warp when that leaves part of the resident wave free for the next launch (programmatic
    // dependent launch) and still gives every SM at least four CTAs (measured on 30000-entry stacks: 625 CTAs = 4.2 per SM);
    // smaller stacks keep the one-wave split, which spreads them over as many warps as possible.  Only in chain mode: a kernel
    // that waits for its predecessor gains nothing from free CTA slots and wants every slot filled (23^3, 30000 entries, waiting
    // mode: 37.6 us with the one-wave split, 49.0 us with 12-entry chunks; profiles/r02_kbench_chain_vs_wait.txt)
    max_chunk = 0;
    if (chained && Policy<M, N, K>::CHUNK > 0 && (long long)Policy<M, N, K>::CHUNK * max_grid * WPC > stack_size &&
        4LL * Policy<M, N, K>::CHUNK * g_num_sms * WPC <= stack_size)
      max_chunk = Policy<M, N, K>::CHUNK;
  }
  if (max_chunk > 0)
    grid = (stack_size + WPC * max_chunk - 1) / (WPC * max_chunk);  // may exceed one resident wave
  else if (grid > max_grid)
    grid = max_grid;
  const int warps = grid * WPC;
  int chunk = (stack_size + warps - 1) / warps, extra = -1;
  if (g_tune.balance.load(std::memory_order_relaxed) > 0) {
    chunk = stack_size / warps;
    extra = stack_size % warps;
  }
  const int align = g_tune.align.load(std::memory_order_relaxed);
  const int flags = ((align < 0 ? Policy<M, N, K>::ALIGN : align != 0) ? FLAG_ALIGN_RUNS : 0) | (chained ? FLAG_PDL_CHAIN : 0);
  unsigned long long* trace = TRACE ? trace_slot() : nullptr;
  const unsigned long long al = a_limit, bl = b_limit;
  return launch_pdl(kern, grid, WPC * 32, SMEM, stream, dev_stack, stack_size, a, b, c, al, bl, chunk, extra, flags, trace);
}

#if defined(SMM_EXPERIMENT)
// warp-specialised kernel (smm_dmma_ws.cuh): NC consumer warps with D stages each + one producer warp per CTA
template <int M, int N, int K, int NC, int D, int HINT, bool TRACE, int FLUSH = 0, bool STAG = false>
int launch_ws(const int* dev_stack, int stack_size, const double* a, const double* b, double* c, uint64_t a_limit, uint64_t b_limit,
              cudaStream_t stream) {
  using G = WsGeom<M, N, K, NC, D>;
  static_assert(G::SMEM <= 227 * 1024, "shared memory budget exceeded");
  auto kern = smm_dmma_ws_kernel<M, N, K, NC, D, HINT, TRACE, FLUSH, STAG>;
  static DevCache cache;
  int g_num_sms = 0;
  const int cps = occupancy(kern, G::THREADS, G::SMEM, cache, g_num_sms);
  if (cps < 0) return cps;
  if (stack_size <= 0) return 0;
  const int max_grid = g_num_sms * cps;
  int grid = (stack_size + NC * 4 - 1) / (NC * 4);
  const int max_chunk = g_tune.chunk.load(std::memory_order_relaxed);
  if (max_chunk > 0)
    grid = (stack_size + NC * max_chunk - 1) / (NC * max_chunk);
  else if (grid > max_grid)
    grid = max_grid;
  // a CTA keeps its slice of the stack (at most stack_size / grid + NC entries) in shared memory
  const int min_grid = (stack_size + (WS_ENT_CAP - NC) - 1) / (WS_ENT_CAP - NC);
  if (grid < min_grid) grid = min_grid;
  const int consumers = grid * NC;
  const int base = stack_size / consumers, extra = stack_size % consumers;
  unsigned long long* trace = TRACE ? trace_slot() : nullptr;
  const unsigned long long al = a_limit, bl = b_limit;
  return launch_pdl(kern, grid, G::THREADS, G::SMEM, stream, dev_stack, stack_size, a, b, c, al, bl, base, extra, trace);
}

#endif

// lane-per-element kernel for tiny blocks (smm_tiny.cuh): no shared memory, WPC warps per CTA, one resident wave
template <int M, int N, int K, int WPC>
int launch_tiny(const int* dev_stack, int stack_size, const double* a, const double* b, double* c, cudaStream_t stream, int policy_chunk = 0) {
  auto kern = smm_tiny_kernel<M, N, K, WPC>;
  static DevCache cache;
  int num_sms = 0;
  const int cps = occupancy(kern, WPC * 32, 0, cache, num_sms);
  if (cps < 0) return cps;
  if (stack_size <= 0) return 0;
  const int max_warps = num_sms * cps * WPC;
  int chunk = g_tune.chunk.load(std::memory_order_relaxed);
  if (chunk < 0) chunk = policy_chunk;  // per-shape policy (autotune database)
  if (chunk <= 0) chunk = (stack_size + max_warps - 1) / max_warps;
  if (chunk < 8) chunk = 8;  // amortise the per-warp prologue and the flush of the last run
  const int warps = (stack_size + chunk - 1) / chunk;
  const int grid = (warps + WPC - 1) / WPC;
  const int flags = (g_use_pdl && stream_chain_mode(stream)) ? FLAG_PDL_CHAIN : 0;
  return launch_pdl(kern, grid, WPC * 32, 0, stream, dev_stack, stack_size, a, b, c, chunk, flags);
}

template <int M, int N, int K>
int launch(const int* dev_stack, int stack_size, const double* a, const double* b, double* c, uint64_t a_limit, uint64_t b_limit,
           cudaStream_t stream) {
  using SH = Shape<M, N, K>;
#if defined(SMM_FORCE_NST) && defined(SMM_FORCE_WPC)
  constexpr int NST = (SMM_FORCE_WPC * SMM_FORCE_NST * SH::STAGE <= 220 * 1024) ? SMM_FORCE_NST : pick_nst(SH::STAGE);
  constexpr int WPC = (SMM_FORCE_WPC * SMM_FORCE_NST * SH::STAGE <= 220 * 1024) ? SMM_FORCE_WPC : pick_wpc(SH::STAGE);
#else
  constexpr int NST = pick_nst(SH::STAGE);
  constexpr int WPC = (Policy<M, N, K>::WPC > 0 && Policy<M, N, K>::WPC * NST * SH::STAGE <= 220 * 1024) ? Policy<M, N, K>::WPC : pick_wpc(SH::STAGE);
#endif
#if defined(SMM_EXPERIMENT)
  // autotune grid carried by EVERY triplet (tools/autotune.sh): 100 + 10 * (index of warps per CTA in {2,4,8}) + FLUSH (0 or 2);
  // 200 / 201 = lane-per-element kernel with 8 / 4 warps per CTA (m*n <= 96 only)
  if constexpr (!(N == M && K == M)) {
#  define SMM_ARGS dev_stack, stack_size, a, b, c, a_limit, b_limit, stream
#  define SMM_TUNE_CASE(ID, WPC_, FLUSH_)                                                                      \
  case ID:                                                                                                    \
    if constexpr (BaseGeom<M, N, K, 1, WPC_, FLUSH_>::SMEM <= 227 * 1024 &&                                    \
                  (FLUSH_ != 2 || Shape<M, N, K>::STAGE >= scratch_bytes(M, N)))                               \
      return launch_base<M, N, K, 1, WPC_, 0, false, FLUSH_>(SMM_ARGS);                                        \
    break;
    switch (g_tune.variant.load(std::memory_order_relaxed)) {
      case 9: return launch_base<M, N, K, NST, WPC, 0, false, 0>(SMM_ARGS);
        SMM_TUNE_CASE(100, 2, 0) SMM_TUNE_CASE(102, 2, 2) SMM_TUNE_CASE(110, 4, 0) SMM_TUNE_CASE(112, 4, 2)
        SMM_TUNE_CASE(120, 8, 0) SMM_TUNE_CASE(122, 8, 2)
      default: break;
    }
#  undef SMM_TUNE_CASE
#  undef SMM_ARGS
  }
  if constexpr (M * N <= TINY_MAX_MN) {
    switch (g_tune.variant.load(std::memory_order_relaxed)) {
      case 200: return launch_tiny<M, N, K, 8>(dev_stack, stack_size, a, b, c, stream);
      case 201: return launch_tiny<M, N, K, 4>(dev_stack, stack_size, a, b, c, stream);
      default: break;
    }
  }
#endif
#if defined(SMM_EXPERIMENT)
  // kernel variants for tools/kbench (one experiment library, run-time switch); only the cubic shape carries them
  if constexpr (N == M && K == M) {
#  define SMM_ARGS dev_stack, stack_size, a, b, c, a_limit, b_limit, stream
    switch (g_tune.variant.load(std::memory_order_relaxed)) {
      case 9: return launch_base<M, N, K, NST, WPC, 0, false, 0>(SMM_ARGS);  // RED flush whatever the policy says
      case 1: return launch_base<M, N, K, NST, WPC, 1, false>(SMM_ARGS);
      case 2: return launch_base<M, N, K, NST, WPC, 2, false>(SMM_ARGS);
      case 3: return launch_base<M, N, K, NST, WPC, 0, true>(SMM_ARGS);
      case 4: return launch_base<M, N, K, 1, 8, 0, false>(SMM_ARGS);
      case 5: return launch_base<M, N, K, 2, 4, 0, false>(SMM_ARGS);
      case 6:
        if constexpr (4 * 6 * SH::STAGE <= 220 * 1024) return launch_base<M, N, K, 6, 4, 0, false>(SMM_ARGS);
        break;
      case 10:
        if constexpr (WsGeom<M, N, K, 8, 3>::SMEM <= 227 * 1024) return launch_ws<M, N, K, 8, 3, 0, false>(SMM_ARGS);
        break;
      case 11:
        if constexpr (WsGeom<M, N, K, 4, 3>::SMEM <= 113 * 1024) return launch_ws<M, N, K, 4, 3, 0, false>(SMM_ARGS);
        break;
      case 12:
        if constexpr (WsGeom<M, N, K, 12, 2>::SMEM <= 227 * 1024) return launch_ws<M, N, K, 12, 2, 0, false>(SMM_ARGS);
        break;
      case 13:
        if constexpr (WsGeom<M, N, K, 4, 6>::SMEM <= 227 * 1024) return launch_ws<M, N, K, 4, 6, 0, false>(SMM_ARGS);
        break;
      case 14:
        if constexpr (WsGeom<M, N, K, 8, 3>::SMEM <= 227 * 1024) return launch_ws<M, N, K, 8, 3, 2, false>(SMM_ARGS);
        break;
      case 15:
        if constexpr (WsGeom<M, N, K, 8, 3>::SMEM <= 227 * 1024) return launch_ws<M, N, K, 8, 3, 0, true>(SMM_ARGS);
        break;
      case 16:
        if constexpr (WsGeom<M, N, K, 6, 4>::SMEM <= 227 * 1024) return launch_ws<M, N, K, 6, 4, 0, false>(SMM_ARGS);
        break;
      case 17:
        if constexpr (WsGeom<M, N, K, 4, 3>::SMEM <= 113 * 1024) return launch_ws<M, N, K, 4, 3, 2, false>(SMM_ARGS);
        break;
      case 18:
        if constexpr (WsGeom<M, N, K, 2, 4>::SMEM <= 75 * 1024) return launch_ws<M, N, K, 2, 4, 0, false>(SMM_ARGS);
        break;
      case 30: return launch_base<M, N, K, 1, 4, 0, false, 1>(SMM_ARGS);
      case 31: return launch_base<M, N, K, 1, 8, 0, false, 1>(SMM_ARGS);
      case 32: return launch_base<M, N, K, 1, 4, 0, true, 1>(SMM_ARGS);
      case 33:
        if constexpr (BaseGeom<M, N, K, 2, 4, 1>::SMEM <= 113 * 1024) return launch_base<M, N, K, 2, 4, 0, false, 1>(SMM_ARGS);
        break;
      case 60: return launch_base<M, N, K, 1, 4, 0, false, 2>(SMM_ARGS);
      case 61: return launch_base<M, N, K, 1, 8, 0, false, 2>(SMM_ARGS);
      case 62: return launch_base<M, N, K, 1, 4, 0, true, 2>(SMM_ARGS);
      case 63: return launch_base<M, N, K, 1, 12, 0, false, 2>(SMM_ARGS);
      case 64:
        if constexpr (BaseGeom<M, N, K, 1, 16, 1>::SMEM <= 227 * 1024) return launch_base<M, N, K, 1, 16, 0, false, 1>(SMM_ARGS);
        break;
      case 70:
        if constexpr (Shape<M, N, K>::ABUF >= scratch_bytes(M, N)) return launch_base<M, N, K, 1, 4, 0, false, 3>(SMM_ARGS);
        break;
      case 71:
        if constexpr (Shape<M, N, K>::ABUF >= scratch_bytes(M, N)) return launch_base<M, N, K, 1, 8, 0, false, 3>(SMM_ARGS);
        break;
      case 72:
        if constexpr (Shape<M, N, K>::ABUF >= scratch_bytes(M, N)) return launch_base<M, N, K, 1, 4, 0, true, 3>(SMM_ARGS);
        break;
      // autotune grid for every cubic shape: 100 + 10 * (index of warps per CTA in {2,4,8,12,16}) + FLUSH (0 or 2), one stage
#  define SMM_TUNE_CASE(ID, WPC_, FLUSH_)                                                                      \
  case ID:                                                                                                    \
    if constexpr (BaseGeom<M, N, K, 1, WPC_, FLUSH_>::SMEM <= 227 * 1024 &&                                    \
                  (FLUSH_ != 2 || Shape<M, N, K>::STAGE >= scratch_bytes(M, N)))                               \
      return launch_base<M, N, K, 1, WPC_, 0, false, FLUSH_>(SMM_ARGS);                                        \
    break;
        SMM_TUNE_CASE(100, 2, 0) SMM_TUNE_CASE(102, 2, 2) SMM_TUNE_CASE(110, 4, 0) SMM_TUNE_CASE(112, 4, 2)
        SMM_TUNE_CASE(120, 8, 0) SMM_TUNE_CASE(122, 8, 2) SMM_TUNE_CASE(130, 12, 0) SMM_TUNE_CASE(132, 12, 2)
        SMM_TUNE_CASE(140, 16, 0) SMM_TUNE_CASE(142, 16, 2)
#  undef SMM_TUNE_CASE
      // deeper per-warp rings for the small shapes (shared memory is not the limit there): 150 + 2 * (NST in {2,4}) + (WPC == 8)
#  define SMM_RING_CASE(ID, NST_, WPC_)                                                                        \
  case ID:                                                                                                    \
    if constexpr (BaseGeom<M, N, K, NST_, WPC_, 0>::SMEM <= 113 * 1024) return launch_base<M, N, K, NST_, WPC_, 0, false, 0>(SMM_ARGS); \
    break;
        SMM_RING_CASE(154, 2, 4) SMM_RING_CASE(155, 2, 8) SMM_RING_CASE(158, 4, 4) SMM_RING_CASE(159, 4, 8)
#  undef SMM_RING_CASE
      case 75: return launch_base<M, N, K, 1, 4, 0, false, 4>(SMM_ARGS);
      case 76: return launch_base<M, N, K, 1, 8, 0, false, 4>(SMM_ARGS);
      case 40: return launch_base<M, N, K, NST, WPC, 0, false, 0, ABL_NOFLUSH>(SMM_ARGS);
      case 41: return launch_base<M, N, K, NST, WPC, 0, false, 0, ABL_NOLDS>(SMM_ARGS);
      case 42: return launch_base<M, N, K, NST, WPC, 0, false, 0, ABL_NOTMA>(SMM_ARGS);
      case 43: return launch_base<M, N, K, NST, WPC, 0, false, 0, ABL_NOFLUSH | ABL_NOLDS | ABL_NOTMA>(SMM_ARGS);
      case 50:
        if constexpr (WsGeom<M, N, K, 8, 3>::SMEM <= 227 * 1024) return launch_ws<M, N, K, 8, 3, 0, false, 2, false>(SMM_ARGS);
        break;
      case 51:
        if constexpr (WsGeom<M, N, K, 8, 3>::SMEM <= 227 * 1024) return launch_ws<M, N, K, 8, 3, 0, false, 2, true>(SMM_ARGS);
        break;
      case 52:
        if constexpr (WsGeom<M, N, K, 4, 3>::SMEM <= 113 * 1024) return launch_ws<M, N, K, 4, 3, 0, false, 2, false>(SMM_ARGS);
        break;
      case 53:
        if constexpr (WsGeom<M, N, K, 4, 3>::SMEM <= 113 * 1024) return launch_ws<M, N, K, 4, 3, 0, false, 2, true>(SMM_ARGS);
        break;
      case 54:
        if constexpr (WsGeom<M, N, K, 12, 2>::SMEM <= 227 * 1024) return launch_ws<M, N, K, 12, 2, 0, false, 2, false>(SMM_ARGS);
        break;
      case 55:
        if constexpr (WsGeom<M, N, K, 12, 2>::SMEM <= 227 * 1024) return launch_ws<M, N, K, 12, 2, 0, false, 2, true>(SMM_ARGS);
        break;
      case 56:
        if constexpr (WsGeom<M, N, K, 8, 3>::SMEM <= 227 * 1024) return launch_ws<M, N, K, 8, 3, 0, false, 0, true>(SMM_ARGS);
        break;
      case 57:
        if constexpr (WsGeom<M, N, K, 8, 3>::SMEM <= 227 * 1024) return launch_ws<M, N, K, 8, 3, 0, true, 2, true>(SMM_ARGS);
        break;
      case 58:
        if constexpr (WsGeom<M, N, K, 8, 2>::SMEM <= 227 * 1024) return launch_ws<M, N, K, 8, 2, 0, false, 2, true>(SMM_ARGS);
        break;
      case 59:
        if constexpr (WsGeom<M, N, K, 2, 4>::SMEM <= 75 * 1024) return launch_ws<M, N, K, 2, 4, 0, false, 2, false>(SMM_ARGS);
        break;
      case 20: return launch_ws<M, N, K, WsPick<M, N, K>::NC, WsPick<M, N, K>::D, 0, false>(SMM_ARGS);
      case 21: return launch_ws<M, N, K, WsPick<M, N, K>::NC, WsPick<M, N, K>::D, 2, false>(SMM_ARGS);
      default: break;
    }
#  undef SMM_ARGS
  }
#endif
  if constexpr (Policy<M, N, K>::ALGO == 1 && M * N <= TINY_MAX_MN)
    return launch_tiny<M, N, K, (Policy<M, N, K>::WPC > 0 ? Policy<M, N, K>::WPC : 8)>(dev_stack, stack_size, a, b, c, stream, Policy<M, N, K>::CHUNK);
  else if constexpr (Policy<M, N, K>::FLUSH == 2 && NST == 1 && SH::STAGE >= scratch_bytes(M, N))
    return launch_base<M, N, K, NST, WPC, 0, false, 2>(dev_stack, stack_size, a, b, c, a_limit, b_limit, stream);
  else
    return launch_base<M, N, K, NST, WPC, 0, false, 0>(dev_stack, stack_size, a, b, c, a_limit, b_limit, stream);
}

template <int M, int N>
launch_fn lookup_k(int k) {
  switch (k) {
#define X(KK) \
  case KK: return launch<M, N, KK>;
    SMM_TUNED_SIZES(X)
#undef X
    default: return nullptr;
  }
}

}  // namespace

#define SMM_CAT2(a, b) a##b
#define SMM_CAT(a, b) SMM_CAT2(a, b)

launch_fn SMM_CAT(lookup_m, SMM_M)(int n, int k) {
  switch (n) {
#define X(NN) \
  case NN: return lookup_k<SMM_M, NN>(k);
    SMM_TUNED_SIZES(X)
#undef X
    default: return nullptr;
  }
}

}  // namespace smm

#if defined(SMM_EXPERIMENT_ONLY_THIS_M)
// kernel-variant experiments build only one M (small libraries travel faster to the GPU box); the other tables are empty
namespace smm {
#  if SMM_M != 5
launch_fn lookup_m5(int, int) { return nullptr; }
#  endif
#  if SMM_M != 13
launch_fn lookup_m13(int, int) { return nullptr; }
#  endif
#  if SMM_M != 23
launch_fn lookup_m23(int, int) { return nullptr; }
#  endif
#  if SMM_M != 26
launch_fn lookup_m26(int, int) { return nullptr; }
#  endif
#  if SMM_M != 32
launch_fn lookup_m32(int, int) { return nullptr; }
#  endif
}  // namespace smm
#endif
