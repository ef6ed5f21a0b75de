// dbcsr_b200/csrc/libsmm_api.cu -- the SMM half of the drop-in C ABI (include/dbcsr_acc_libsmm.h).
//
// libsmm_acc_process / libsmm_acc_transpose / c_calculate_norms with the reference's argument meaning and return codes
// (src/acc/libsmm_acc/libsmm_acc.cpp:324-339,482-487; src/acc/cuda_hip/calculate_norms.cpp:98-117), dispatching to the
// ahead-of-time compiled sm_100a kernels of this library.  No JIT, no CPU fall-back inside the library: an unsupported request
// returns a negative code and leaves C untouched (DBCSR then runs the stack on its own CPU driver).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>
#include <vector>

#include "../../include/dbcsr_acc_libsmm.h"
#include "smm_bf16.cuh"
#include "smm_bf16_tiled.cuh"
#include "smm_bf16_plan.cuh"
#include "smm_dmma_big.cuh"
#include "smm_dmma_huge.cuh"
#include "smm_dmma_rt.cuh"
#include "smm_generic.cuh"
#include "smm_launch.h"
#include "smm_tune.h"

namespace smm {
Tunables g_tune;

// ---- chain registry (see smm_launch.h).  A handful of streams at most: linear search under a mutex, skipped when empty.
namespace {
struct ChainEntry {
  cudaStream_t stream;
  bool primed;  // the last launch of this library on the stream was an independent FP64 stack drain
};
std::mutex g_chain_mu;
std::vector<ChainEntry> g_chain;
std::atomic<int> g_chain_n{0};
// DBCSR_B200_PDL_CHAIN=1: every stream is a chain (the caller guarantees that nothing but stack drains, stack uploads and event
// waits is ever enqueued between two libsmm_acc_process calls of a stream)
const bool g_chain_all = [] {
  const char* e = getenv("DBCSR_B200_PDL_CHAIN");
  return e != nullptr && atoi(e) != 0;
}();
}  // namespace

bool stream_chain_mode(cudaStream_t stream) {
  if (g_chain_all) return true;
  if (g_chain_n.load(std::memory_order_acquire) == 0) return false;
  std::lock_guard<std::mutex> lock(g_chain_mu);
  for (auto& c : g_chain)
    if (c.stream == stream) {
      const bool was = c.primed;
      c.primed = true;
      return was;
    }
  return false;
}
void stream_chain_break(cudaStream_t stream) {
  if (g_chain_n.load(std::memory_order_acquire) == 0) return;
  std::lock_guard<std::mutex> lock(g_chain_mu);
  for (auto& c : g_chain)
    if (c.stream == stream) c.primed = false;
}
}  // namespace smm

namespace {

// environment defaults of the run-time knobs (smm_tune.h), read once when the library is loaded
const bool g_tune_env_read = [] {
  if (const char* e = getenv("DBCSR_B200_BALANCE")) smm::g_tune.balance.store(atoi(e));
  if (const char* e = getenv("DBCSR_B200_CHUNK")) smm::g_tune.chunk.store(atoi(e));
  if (const char* e = getenv("DBCSR_B200_ALIGN")) smm::g_tune.align.store(atoi(e));
  if (const char* e = getenv("DBCSR_B200_VARIANT")) smm::g_tune.variant.store(atoi(e));
  if (const char* e = getenv("DBCSR_B200_BIGDMMA")) smm::g_tune.bigdmma.store(atoi(e));
  if (const char* e = getenv("DBCSR_B200_HUGEDMMA")) smm::g_tune.hugedmma.store(atoi(e));
  if (const char* e = getenv("DBCSR_B200_INHOMOGENEOUS")) smm::g_tune.inhomogeneous.store(atoi(e));
  if (const char* e = getenv("DBCSR_B200_BF16_MERGE")) smm::g_tune.bf16_merge.store(atoi(e));
  if (const char* e = getenv("DBCSR_B200_BF16_A_TMEM")) smm::g_tune.bf16_a_tmem.store(atoi(e));
  if (const char* e = getenv("DBCSR_B200_BF16_PLAN")) smm::g_tune.bf16_plan.store(atoi(e));
  return true;
}();

std::atomic<long long> g_launches{0};
std::atomic<int> g_num_sms{0};

int num_sms() {
  int n = g_num_sms.load(std::memory_order_relaxed);
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess)
      g_num_sms.store(n, std::memory_order_relaxed);
    else
      n = 148;
  }
  return n;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: remember the largest value set per device (DBCSR picks
// the device per rank and c_dbcsr_acc_set_active_device may switch it)
struct SmemAttrCache {
  std::atomic<int> set[64];
};
template <typename Kern>
int ensure_smem(Kern kern, int smem, SmemAttrCache& cache) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -30;
  if (cache.set[dev].load(std::memory_order_acquire) < smem) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -30;
    cache.set[dev].store(smem, std::memory_order_release);
  }
  return 0;
}

// End address of the device allocation that contains `p` (0 = unknown).  The DMMA kernel stages 16-byte-aligned windows and may
// over-read up to 8 bytes behind a block; for the last block of an allocation that must not cross the allocation's end.
// cuMemGetAddressRange is fetched through the runtime so that the library has no link-time dependency on libcuda.
typedef int (*cuMemGetAddressRange_t)(unsigned long long* pbase, size_t* psize, unsigned long long dptr);
uint64_t allocation_end(const void* p) {
  static cuMemGetAddressRange_t fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    (void)cudaGetLastError();
    return reinterpret_cast<cuMemGetAddressRange_t>(f);
  }();
  if (fn == nullptr || p == nullptr) return 0;
  unsigned long long base = 0;
  size_t size = 0;
  if (fn(&base, &size, reinterpret_cast<unsigned long long>(p)) != 0) return 0;
  return base + size;
}

smm::launch_fn lookup(int m, int n, int k) {
  switch (m) {
    case 5: return smm::lookup_m5(n, k);
    case 13: return smm::lookup_m13(n, k);
    case 23: return smm::lookup_m23(n, k);
    case 26: return smm::lookup_m26(n, k);
    case 32: return smm::lookup_m32(n, k);
    default: return nullptr;
  }
}

int launch_bf16(const int* dev_stack, int stack_size, const void* a_tiles, const void* b_tiles, void* c, int m, int n, int k,
                cudaStream_t stream) {
  if (stack_size <= 0) return 0;
  const smm::Bf16Geom g = smm::bf16_geom(m, n, k);
  const size_t smem = smm::bf16_smem_bytes(g);
  static SmemAttrCache smem_set;
  if (ensure_smem(smm::smm_bf16_kernel, (int)smem, smem_set) != 0) return -30;
  int per_sm = (int)((220 * 1024) / smem);
  if (per_sm > 512 / smm::BF_TMEM_COLS) per_sm = 512 / smm::BF_TMEM_COLS;  // TMEM: 512 columns per SM
  if (per_sm < 1) return -30;
  const int max_grid = num_sms() * per_sm;
  int grid = (stack_size + 15) / 16;
  if (grid > max_grid) grid = max_grid;
  if (per_sm > 8) per_sm = 8;
  const int chunk = (stack_size + grid - 1) / grid;
  grid = (stack_size + chunk - 1) / chunk;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(smm::BF_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t err = cudaLaunchKernelEx(&cfg, smm::smm_bf16_kernel, dev_stack, stack_size, static_cast<const unsigned char*>(a_tiles),
                                             static_cast<const unsigned char*>(b_tiles), static_cast<float*>(c), m, n, k, chunk);
  return (err == cudaSuccess) ? 0 : -31;
}

// run-time-shape DMMA kernel (smm_dmma_rt.cuh): every m, n <= 32 without a specialised kernel, k limited by shared memory
template <int TM, int TN>
int launch_rt_t(const int* dev_stack, int stack_size, const double* a, const double* b, double* c, int m, int n, int k, uint64_t a_end,
                uint64_t b_end, cudaStream_t stream) {
  const int smem = 128 + smm::RT_WPC * (smm::rt_abuf(m, k) + smm::rt_abuf(n, k));
  static SmemAttrCache smem_set;
  if (ensure_smem(smm::smm_dmma_rt_kernel<TM, TN>, smem, smem_set) != 0) return -30;
  int cps = (220 * 1024) / smem;
  if (cps > 12) cps = 12;
  if (cps < 1) return -30;
  if (stack_size <= 0) return 0;  // stack_size 0 = prepare only
  const int max_grid = num_sms() * cps;
  int grid = (stack_size + smm::RT_WPC * 4 - 1) / (smm::RT_WPC * 4);
  if (grid > max_grid) grid = max_grid;
  const int warps = grid * smm::RT_WPC;
  const int chunk = (stack_size + warps - 1) / warps;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(smm::RT_WPC * 32);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const unsigned long long al = a_end, bl = b_end;
  const cudaError_t err = cudaLaunchKernelEx(&cfg, smm::smm_dmma_rt_kernel<TM, TN>, dev_stack, stack_size, a, b, c, al, bl, chunk, m, n, k);
  return (err == cudaSuccess) ? 0 : -31;
}

// cooperative DMMA kernel (smm_dmma_big.cuh) for blocks with a dimension in 33..80 (default; "bigdmma" = 0 falls back to the generic kernel)
bool big_eligible(int m, int n, int k) {
  return smm::g_tune.bigdmma.load(std::memory_order_relaxed) != 0 && m <= 80 && n <= smm::BIG_MAX_N && k <= 80 && m > 0 && n > 0 && k > 0 &&
         smm::big_smem_bytes(m, n, k) <= 110 * 1024;
}

int launch_big(const int* dev_stack, int stack_size, const double* a, const double* b, double* c, int m, int n, int k, uint64_t a_end,
               uint64_t b_end, cudaStream_t stream) {
  const int smem = smm::big_smem_bytes(m, n, k);
  static SmemAttrCache smem_set;
  if (ensure_smem(smm::smm_dmma_big_kernel, smem, smem_set) != 0) return -30;
  int cps = (220 * 1024) / (smem + 1024);
  if (cps > 8) cps = 8;
  if (cps < 1) return -30;
  if (stack_size <= 0) return 0;  // stack_size 0 = prepare only
  const int max_grid = num_sms() * cps;
  int grid = (stack_size + 1) / 2;  // at least two entries per CTA
  if (grid > max_grid) grid = max_grid;
  const int chunk = (stack_size + grid - 1) / grid;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(smm::BIG_WARPS * 32);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const unsigned long long al = a_end, bl = b_end;
  const cudaError_t err = cudaLaunchKernelEx(&cfg, smm::smm_dmma_big_kernel, dev_stack, stack_size, a, b, c, al, bl, chunk, m, n, k);
  return (err == cudaSuccess) ? 0 : -31;
}

bool rt_eligible(int m, int n, int k) {
  return m <= 32 && n <= 32 && 128 + smm::RT_WPC * (smm::rt_abuf(m, k) + smm::rt_abuf(n, k)) <= 220 * 1024;
}

int launch_rt(const int* dev_stack, int stack_size, const double* a, const double* b, double* c, int m, int n, int k, uint64_t a_end,
              uint64_t b_end, cudaStream_t stream) {
  const int tm = (m + 7) / 8, tn = (n + 7) / 8;
#define SMM_RT_CASE(TM_, TN_) \
  if (tm == TM_ && tn == TN_) return launch_rt_t<TM_, TN_>(dev_stack, stack_size, a, b, c, m, n, k, a_end, b_end, stream);
  SMM_RT_CASE(1, 1) SMM_RT_CASE(1, 2) SMM_RT_CASE(1, 3) SMM_RT_CASE(1, 4)
  SMM_RT_CASE(2, 1) SMM_RT_CASE(2, 2) SMM_RT_CASE(2, 3) SMM_RT_CASE(2, 4)
  SMM_RT_CASE(3, 1) SMM_RT_CASE(3, 2) SMM_RT_CASE(3, 3) SMM_RT_CASE(3, 4)
  SMM_RT_CASE(4, 1) SMM_RT_CASE(4, 2) SMM_RT_CASE(4, 3) SMM_RT_CASE(4, 4)
#undef SMM_RT_CASE
  return -30;
}

// blocks above max_kernel_dim: panel DMMA kernel (smm_dmma_huge.cuh), one CTA per entry in turn
int launch_huge(const int* dev_stack, int stack_size, const double* a, const double* b, double* c, int m, int n, int k, int b_transposed,
                cudaStream_t stream) {
  if (stack_size <= 0) return 0;
  int grid = stack_size;
  const int max_grid = num_sms() * 4;
  if (grid > max_grid) grid = max_grid;
  smm::smm_dmma_huge_kernel<<<grid, smm::BIG_WARPS * 32, 0, stream>>>(dev_stack, stack_size, a, b, c, m, n, k, b_transposed);
  return (cudaPeekAtLastError() == cudaSuccess) ? 0 : -31;
}

int launch_generic(const int* dev_stack, int stack_size, const double* a, const double* b, double* c, int m, int n, int k,
                   int b_transposed, cudaStream_t stream) {
  if (stack_size <= 0) return 0;
  // blocks with a dimension above 80 (the only FP64 shapes that reach this function besides bigdmma = 0): tensor-pipe panel kernel
  if ((m > 80 || n > 80 || k > 80) && smm::g_tune.hugedmma.load(std::memory_order_relaxed) != 0)
    return launch_huge(dev_stack, stack_size, a, b, c, m, n, k, b_transposed, stream);
  const int max_grid = num_sms() * 8;
  int grid = (stack_size + smm::GEN_WPC * 2 - 1) / (smm::GEN_WPC * 2);
  if (grid > max_grid) grid = max_grid;
  const int warps = grid * smm::GEN_WPC;
  const int chunk = (stack_size + warps - 1) / warps;
  smm::smm_generic_kernel<<<grid, smm::GEN_WPC * 32, 0, stream>>>(dev_stack, stack_size, a, b, c, m, n, k, b_transposed, chunk);
  return (cudaPeekAtLastError() == cudaSuccess) ? 0 : -31;
}

// ---- inhomogeneous stacks (def_mnk = 0) -------------------------------------------------------------------------------------
// The reference rejects them (-1, libsmm_acc.cpp:327) and DBCSR drains them on the CPU.  Here the 7-wide HOST stack is binned by
// (m,n,k), every bin ordered by c_first (stable), uploaded through a small ring of pinned/device scratch buffers per stream and
// drained by the kernel of its shape, all on stack_stream.  The host stack is consumed before the call returns.
// All-or-nothing: every bin's kernel is resolved and prepared (shared-memory attribute, occupancy) BEFORE anything is enqueued,
// so a negative return code ("redo this stack on the CPU", src/mm/dbcsr_acc_operations.F:134-135) always means C is untouched.
struct InhomoScratch {
  int* host = nullptr;
  int* dev = nullptr;
  size_t cap = 0;  // ints
  cudaEvent_t done = nullptr;
};
struct InhomoRing {
  cudaStream_t stream = nullptr;
  InhomoScratch buf[4];
  int next = 0;
};
// Rings live as long as the library (host threads of the engine are short-lived; freeing pinned/device memory at thread exit
// would synchronise the device under in-flight work): one per stream, created on first use, released by libsmm_acc_finalize.
std::mutex g_inhomo_mu;
std::vector<InhomoRing*> g_inhomo_rings;

InhomoRing* inhomo_ring(cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(g_inhomo_mu);
  for (InhomoRing* r : g_inhomo_rings)
    if (r->stream == stream) return r;
  InhomoRing* r = new InhomoRing();
  r->stream = stream;
  g_inhomo_rings.push_back(r);
  return r;
}
void inhomo_release_all() {
  std::lock_guard<std::mutex> lock(g_inhomo_mu);
  for (InhomoRing* r : g_inhomo_rings) {
    for (auto& b : r->buf) {
      if (b.done != nullptr) {
        cudaEventSynchronize(b.done);
        cudaEventDestroy(b.done);
      }
      if (b.host != nullptr) cudaFreeHost(b.host);
      if (b.dev != nullptr) cudaFree(b.dev);
    }
    delete r;
  }
  g_inhomo_rings.clear();
}

enum BinKind { BIN_EMPTY, BIN_TUNED, BIN_RT, BIN_BIG, BIN_GENERIC_BT, BIN_GENERIC_NT };
struct Bin {
  int lo, hi, m, n, k;
  BinKind kind;
  smm::launch_fn fn;
};

int process_inhomogeneous(const int* host7, int stack_size, const double* a, const double* b, double* c, int max_kernel_dim,
                          cudaStream_t stream) {
  if (host7 == nullptr) return -1;
  if (stack_size <= 0) return 0;
  // order: by shape, then by c_first, stable
  std::vector<int> order((size_t)stack_size);
  std::iota(order.begin(), order.end(), 0);
  auto shape = [&](int i) {
    const int* p = host7 + 7 * (size_t)i;
    return ((uint64_t)(uint32_t)p[0] << 42) | ((uint64_t)(uint32_t)p[1] << 21) | (uint64_t)(uint32_t)p[2];
  };
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
    const uint64_t sx = shape(x), sy = shape(y);
    if (sx != sy) return sx < sy;
    return host7[7 * (size_t)x + 5] < host7[7 * (size_t)y + 5];
  });
  // ---- phase 1: resolve and prepare every bin; nothing is enqueued yet, any failure leaves C untouched
  const uint64_t a_end = allocation_end(a), b_end = allocation_end(b);
  std::vector<Bin> bins;
  for (int lo = 0; lo < stack_size;) {
    int hi = lo;
    const uint64_t sh = shape(order[(size_t)lo]);
    while (hi < stack_size && shape(order[(size_t)hi]) == sh) ++hi;
    const int* p = host7 + 7 * (size_t)order[(size_t)lo];
    Bin bin = {lo, hi, p[0], p[1], p[2], BIN_EMPTY, nullptr};
    const int m = bin.m, n = bin.n, k = bin.k;
    int rc = 0;
    if (m <= 0 || n <= 0 || k <= 0) {
      bin.kind = BIN_EMPTY;  // empty blocks: nothing to do
    }
    else if (m > max_kernel_dim || n > max_kernel_dim || k > max_kernel_dim) {
      bin.kind = (n <= max_kernel_dim && k <= max_kernel_dim) ? BIN_GENERIC_BT : BIN_GENERIC_NT;
    }
    else if ((bin.fn = lookup(m, n, k)) != nullptr) {
      bin.kind = BIN_TUNED;
      rc = bin.fn(nullptr, 0, a, b, c, a_end, b_end, stream);  // stack_size 0 = prepare only
    }
    else if (rt_eligible(m, n, k)) {
      bin.kind = BIN_RT;
      rc = launch_rt(nullptr, 0, a, b, c, m, n, k, a_end, b_end, stream);
    }
    else if (big_eligible(m, n, k)) {
      bin.kind = BIN_BIG;
      rc = launch_big(nullptr, 0, a, b, c, m, n, k, a_end, b_end, stream);
    }
    else {
      bin.kind = BIN_GENERIC_BT;
    }
    if (rc < 0) return rc;
    bins.push_back(bin);
    lo = hi;
  }
  InhomoRing* ring = inhomo_ring(stream);  // calls on one stream come from one host thread at a time (DBCSR: one stream per thread)
  InhomoScratch& sc = ring->buf[ring->next];
  ring->next = (ring->next + 1) % 4;
  if (sc.done == nullptr && cudaEventCreateWithFlags(&sc.done, cudaEventDisableTiming) != cudaSuccess) return -30;
  if (cudaEventSynchronize(sc.done) != cudaSuccess) return -30;  // previous user of this scratch has finished
  const size_t need = 3 * (size_t)stack_size;
  if (sc.cap < need) {
    if (sc.host != nullptr) cudaFreeHost(sc.host);
    if (sc.dev != nullptr) cudaFree(sc.dev);
    sc.host = nullptr;
    sc.dev = nullptr;
    sc.cap = 0;
    const size_t cap = std::max(need, (size_t)3 * 30000);
    if (cudaHostAlloc(reinterpret_cast<void**>(&sc.host), cap * sizeof(int), cudaHostAllocDefault) != cudaSuccess) {
      sc.host = nullptr;
      (void)cudaGetLastError();
      return -30;
    }
    if (cudaMalloc(reinterpret_cast<void**>(&sc.dev), cap * sizeof(int)) != cudaSuccess) {
      cudaFreeHost(sc.host);
      sc.host = nullptr;
      sc.dev = nullptr;
      (void)cudaGetLastError();
      return -30;
    }
    sc.cap = cap;  // only now: both buffers exist
  }
  for (int i = 0; i < stack_size; ++i) {
    const int* p = host7 + 7 * (size_t)order[(size_t)i];
    sc.host[3 * (size_t)i] = p[3];
    sc.host[3 * (size_t)i + 1] = p[4];
    sc.host[3 * (size_t)i + 2] = p[5];
  }
  if (cudaMemcpyAsync(sc.dev, sc.host, need * sizeof(int), cudaMemcpyHostToDevice, stream) != cudaSuccess) return -31;  // nothing enqueued
  // ---- phase 2: launch.  Every kernel was prepared above, so a failure here is a device/driver fault, not an unsupported
  // request: like the reference (ACC_API_CALL: print and exit, src/acc/cuda/acc_cuda.h:29-36) it is fatal -- a negative code
  // would make DBCSR redo the WHOLE stack on the CPU on top of the bins already enqueued.
  smm::stream_chain_break(stream);  // the bins read a stack uploaded by this very call
  int rc_all = 0;
  for (const Bin& bin : bins) {
    const int* st = sc.dev + 3 * (size_t)bin.lo;
    const int cnt = bin.hi - bin.lo;
    int rc = 0;
    switch (bin.kind) {
      case BIN_EMPTY: continue;
      case BIN_TUNED: rc = bin.fn(st, cnt, a, b, c, a_end, b_end, stream); break;
      case BIN_RT: rc = launch_rt(st, cnt, a, b, c, bin.m, bin.n, bin.k, a_end, b_end, stream); break;
      case BIN_BIG: rc = launch_big(st, cnt, a, b, c, bin.m, bin.n, bin.k, a_end, b_end, stream); break;
      case BIN_GENERIC_BT: rc = launch_generic(st, cnt, a, b, c, bin.m, bin.n, bin.k, 1, stream); break;
      case BIN_GENERIC_NT: rc = launch_generic(st, cnt, a, b, c, bin.m, bin.n, bin.k, 0, stream); break;
    }
    if (rc < 0) {
      fprintf(stderr, "dbcsr_acc_b200: kernel launch failed (%d: %s) for the %dx%dx%d bin of an inhomogeneous stack after earlier bins were enqueued\n",
              rc, cudaGetErrorString(cudaGetLastError()), bin.m, bin.n, bin.k);
      abort();
    }
    if (bin.kind != BIN_TUNED) rc_all = 10;
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  if (cudaEventRecord(sc.done, stream) != cudaSuccess) {
    fprintf(stderr, "dbcsr_acc_b200: cudaEventRecord failed behind an inhomogeneous stack\n");
    abort();
  }
  return rc_all;
}

template <typename T>
int launch_generic_typed(const int* dev_stack, int stack_size, const void* a, const void* b, void* c, int m, int n, int k,
                         int b_transposed, cudaStream_t stream) {
  if (stack_size <= 0) return 0;
  const int max_grid = num_sms() * 8;
  int grid = (stack_size + 15) / 16;
  if (grid > max_grid) grid = max_grid;
  const int warps = grid * 8;
  const int chunk = (stack_size + warps - 1) / warps;
  smm::smm_generic_typed_kernel<T><<<grid, 256, 0, stream>>>(dev_stack, stack_size, static_cast<const T*>(a), static_cast<const T*>(b),
                                                            static_cast<T*>(c), m, n, k, b_transposed, chunk);
  return (cudaPeekAtLastError() == cudaSuccess) ? 0 : -31;
}

template <typename T>
int launch_transpose_typed(const int* dev_trs_stack, int nblks, void* data, int m, int n, cudaStream_t stream) {
  const size_t blk_bytes = (size_t)m * n * sizeof(T);
  int wpc = (int)((96 * 1024) / blk_bytes);
  if (wpc > 8) wpc = 8;
  if (wpc < 1) wpc = 1;  // 80 x 80 complex_8 = 100 KB: one warp per CTA, still within the 227 KB of an SM
  if (blk_bytes * (size_t)wpc > 200 * 1024) return -3;  // cannot happen for m,n <= 80
  static SmemAttrCache attr_set;
  if (ensure_smem(smm::transpose_typed_kernel<T>, (int)(blk_bytes * wpc), attr_set) != 0) return -30;
  int grid = (nblks + wpc - 1) / wpc;
  const int max_grid = num_sms() * 8;
  if (grid > max_grid) grid = max_grid;
  smm::transpose_typed_kernel<T><<<grid, wpc * 32, blk_bytes * wpc, stream>>>(dev_trs_stack, nblks, static_cast<T*>(data), m, n);
  return (cudaPeekAtLastError() == cudaSuccess) ? 0 : -31;
}

// DBCSR_B200_ALL_TYPES=0 restores the reference's behaviour for real_4 / complex types (process: -10, transpose: no-op)
bool all_types_enabled() {
  static const bool v = [] {
    const char* e = getenv("DBCSR_B200_ALL_TYPES");
    return e == nullptr || atoi(e) != 0;
  }();
  return v;
}

// Zeroing at a bounded rate: `gridDim.x` CTAs (a handful) stream 16-byte zero stores over the range.  cudaMemsetAsync zeroes a
// 4 GB buffer in under a millisecond -- and for that millisecond takes most of the HBM bandwidth away from stack kernels running
// beside it; a caller that zeroes the NEXT multiply's C buffer while this multiply's stacks run (bench.py, cannon replay) wants the
// same bytes spread over the whole multiply instead.
__global__ void __launch_bounds__(1024) zero_trickle_kernel(uint4* __restrict__ p, size_t n16) {
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) __stcs(p + i, z);
}

// Register-only DMMA.8x8x4 loop: the FP64 tensor-pipe peak of THIS device, measured in place so that roofline fractions have a
// live denominator (bench.py).  9 independent accumulator pairs per warp like the 23^3 kernel's 3 x 3 tiles.  The operands carry
// RANDOM mantissas (per-thread hash); the probe holds its rate for as long as it runs (the pipe alone does not reach the board's
// power limit, see below).
// FRESH (experiment, not exported): new operand mantissas every iteration.  Measured on B200: 32.8 TFLOP/s burst AND sustained -- the
// extra integer work costs issue slots, and the power limit is still not reached; cuBLAS DGEMM 8192^3 on uniform(0,1) data does not
// throttle either (35.5 TFLOP/s for 0.4 s).  What pulls the clock down under the stack kernel is its memory traffic (44 GB of DRAM and
// ~110 GB of L2 -> SM traffic per 10 ms multiply), not the DMMA operand bits: on the device builder's tile-ordered stacks (a third of
// the DRAM traffic) the same kernel holds its burst rate much longer.
template <bool FRESH>
__global__ void __launch_bounds__(512) fp64_peak_kernel(double* out, int iters, unsigned seed) {
  double c0[9], c1[9];
  unsigned long long h = (unsigned long long)(blockIdx.x * blockDim.x + threadIdx.x + 1) * 0x9E3779B97F4A7C15ull + seed;
  auto rnd = [&]() {  // uniform in [0,1) with 52 random mantissa bits
    h ^= h >> 12;
    h ^= h << 25;
    h ^= h >> 27;
    return __longlong_as_double((long long)(((h * 2685821657736338717ull) >> 12) | 0x3FF0000000000000ull)) - 1.0;
  };
#pragma unroll
  for (int i = 0; i < 9; ++i) c0[i] = rnd(), c1[i] = rnd();
  double a[3], b[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) a[i] = rnd(), b[i] = rnd() - 0.5;
  for (int it = 0; it < iters; ++it) {
    if (FRESH) {
      h ^= h >> 12;
      h ^= h << 25;
      h ^= h >> 27;
#pragma unroll
      for (int i = 0; i < 3; ++i) {  // new mantissa bits, exponent and sign kept: values stay in their range
        a[i] = __longlong_as_double(__double_as_longlong(a[i]) ^ (long long)((h >> (4 * i)) & 0x000FFFFFFFFFFFF0ull));
        b[i] = __longlong_as_double(__double_as_longlong(b[i]) ^ (long long)((h >> (4 * i + 2)) & 0x0007FFFFFFFFFFF0ull));
      }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) smm::dmma884(c0[3 * i + j], c1[3 * i + j], a[i], b[j]);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) s += c0[i] + c1[i];
  if (s == 1.2345e300) out[0] = s;
}


// plan buffers of the tiled BF16 SpGEMM, one per stream, kept for the lifetime of the library (a multiply re-derives its plan
// on the stream it runs on, so consecutive calls on one stream may share the buffer; `zeros_off`: the zero tile, cleared once)
struct BtScratch {
  cudaStream_t st;
  int dev;
  unsigned char* buf;
  size_t cap;
};
std::mutex g_bt_mu;
std::vector<BtScratch> g_bt_scratch;
unsigned char* bt_plan_scratch(cudaStream_t st, size_t bytes, size_t zeros_off) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_bt_mu);
  BtScratch* e = nullptr;
  for (auto& x : g_bt_scratch)
    if (x.st == st && x.dev == dev) e = &x;
  if (e == nullptr) {
    g_bt_scratch.push_back(BtScratch{st, dev, nullptr, 0});
    e = &g_bt_scratch.back();
  }
  if (e->cap < bytes) {
    if (e->buf != nullptr) {
      cudaStreamSynchronize(st);
      cudaFree(e->buf);
      e->buf = nullptr;
      e->cap = 0;
    }
    if (cudaMalloc(reinterpret_cast<void**>(&e->buf), bytes) != cudaSuccess) return nullptr;
    e->cap = bytes;
  }
  // the zero tile sits at a size-dependent offset: clear it on every call (10 KB)
  if (cudaMemsetAsync(e->buf + zeros_off, 0, 5 * 2048, st) != cudaSuccess) return nullptr;
  return e->buf;
}
}  // namespace

extern "C" {

int libsmm_acc_init(void) { return 0; }      // kernels are compiled ahead of time; nothing to set up per thread
int libsmm_acc_finalize(void) {
  inhomo_release_all();
  return 0;
}
c_dbcsr_acc_bool_t libsmm_acc_is_thread_safe(void) { return 1; }
int libsmm_acc_gpu_warp_size(void) { return 32; }
long long libsmm_acc_b200_launch_count(void) { return g_launches.load(); }
const char* libsmm_acc_b200_version(void) { return "dbcsr_acc_b200 r1 (sm_100a, DMMA.8x8x4 + TMA bulk staging)"; }

// Run-time knobs of the FP64 stack kernels (smm_tune.h).  Names: "balance", "align", "chunk", "bigdmma", "variant", "trace_first", "trace_count",
// "seq" (launch sequence counter).  Returns 0, or -1 for an unknown name.  libsmm_acc_b200_set_trace installs a device buffer of
// trace_count * 4096 * 128 64-bit words (NULL switches tracing off); only TRACE kernel variants of experiment builds write to it.
int libsmm_acc_b200_set_tunable(const char* name, long long value) {
  if (name == nullptr) return -1;
  if (strcmp(name, "balance") == 0) smm::g_tune.balance.store((int)value);
  else if (strcmp(name, "chunk") == 0) smm::g_tune.chunk.store((int)value);
  else if (strcmp(name, "align") == 0) smm::g_tune.align.store((int)value);
  else if (strcmp(name, "variant") == 0) smm::g_tune.variant.store((int)value);
  else if (strcmp(name, "bigdmma") == 0) smm::g_tune.bigdmma.store((int)value);
  else if (strcmp(name, "hugedmma") == 0) smm::g_tune.hugedmma.store((int)value);
  else if (strcmp(name, "inhomogeneous") == 0) smm::g_tune.inhomogeneous.store((int)value);
  else if (strcmp(name, "bf16_merge") == 0) smm::g_tune.bf16_merge.store((int)value);
  else if (strcmp(name, "bf16_a_tmem") == 0) smm::g_tune.bf16_a_tmem.store((int)value);
  else if (strcmp(name, "bf16_plan") == 0) smm::g_tune.bf16_plan.store((int)value);
  else if (strcmp(name, "trace_first") == 0) smm::g_tune.trace_first.store((int)value);
  else if (strcmp(name, "trace_count") == 0) smm::g_tune.trace_count.store((int)value);
  else if (strcmp(name, "seq") == 0) smm::g_tune.seq.store((int)value);
  else return -1;
  return 0;
}
long long libsmm_acc_b200_get_tunable(const char* name) {
  if (name == nullptr) return -1;
  if (strcmp(name, "balance") == 0) return smm::g_tune.balance.load();
  if (strcmp(name, "chunk") == 0) return smm::g_tune.chunk.load();
  if (strcmp(name, "align") == 0) return smm::g_tune.align.load();
  if (strcmp(name, "variant") == 0) return smm::g_tune.variant.load();
  if (strcmp(name, "bigdmma") == 0) return smm::g_tune.bigdmma.load();
  if (strcmp(name, "hugedmma") == 0) return smm::g_tune.hugedmma.load();
  if (strcmp(name, "inhomogeneous") == 0) return smm::g_tune.inhomogeneous.load();
  if (strcmp(name, "bf16_merge") == 0) return smm::g_tune.bf16_merge.load();
  if (strcmp(name, "bf16_a_tmem") == 0) return smm::g_tune.bf16_a_tmem.load();
  if (strcmp(name, "bf16_plan") == 0) return smm::g_tune.bf16_plan.load();
  if (strcmp(name, "trace_first") == 0) return smm::g_tune.trace_first.load();
  if (strcmp(name, "trace_count") == 0) return smm::g_tune.trace_count.load();
  if (strcmp(name, "seq") == 0) return smm::g_tune.seq.load();
  if (strcmp(name, "experiment") == 0) {
#if defined(SMM_EXPERIMENT)
    return 1;
#else
    return 0;
#endif
  }
  return -1;
}
// Declares (on != 0) or withdraws (on == 0) that `stream` carries a CHAIN of independent stack drains: between two
// libsmm_acc_process calls on it the caller enqueues nothing that produces A, B, C or stack data except through operations that
// are full stream dependencies anyway (memcpys, event waits).  The FP64 stack kernels then skip the grid-dependency wait in
// front of their first read, so consecutive drains overlap tail and ramp-up (programmatic dependent launch).  The first drain
// after the declaration -- and after any other kernel this library launches on the stream -- still waits.  Default: off.
int libsmm_acc_b200_stream_chain(void* stream, int on) {
  if (stream == nullptr) return -2;
  const cudaStream_t st = *static_cast<cudaStream_t*>(stream);
  std::lock_guard<std::mutex> lock(smm::g_chain_mu);
  for (size_t i = 0; i < smm::g_chain.size(); ++i)
    if (smm::g_chain[i].stream == st) {
      if (on == 0) {
        smm::g_chain.erase(smm::g_chain.begin() + (long)i);
        smm::g_chain_n.store((int)smm::g_chain.size(), std::memory_order_release);
      }
      else {
        smm::g_chain[i].primed = false;
      }
      return 0;
    }
  if (on != 0) {
    smm::g_chain.push_back({st, false});
    smm::g_chain_n.store((int)smm::g_chain.size(), std::memory_order_release);
  }
  return 0;
}

void libsmm_acc_b200_set_trace(void* dev_words) { smm::g_tune.trace.store(static_cast<unsigned long long*>(dev_words)); }

// memset_zero at a bounded rate (see zero_trickle_kernel): `nctas` CTAs of 1024 threads; offset and nbytes multiples of 16.
int libsmm_acc_b200_memset_zero_trickle(void* dev_mem, size_t offset, size_t nbytes, int nctas, void* stream) {
  if (nbytes == 0) return 0;
  if (dev_mem == nullptr || stream == nullptr || nctas < 1 || (offset & 15) != 0 || (nbytes & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(dev_mem) & 15) != 0)
    return -2;
  zero_trickle_kernel<<<nctas, 1024, 0, *static_cast<cudaStream_t*>(stream)>>>(reinterpret_cast<uint4*>(static_cast<char*>(dev_mem) + offset),
                                                                            nbytes / 16);
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -31;
}

// Measured FP64 tensor-pipe (DMMA.8x8x4) throughput of the active device in GFLOP/s: 16 warps per SM, register operands, best of
// three timed launches on `stream` (synchronises it).  Returns <= 0 on failure.  Introspection only: not on any product path.
static void launch_peak(int sms, int warps, int iters, bool fresh, cudaStream_t st) {
  if (fresh)
    fp64_peak_kernel<true><<<sms, warps * 32, 0, st>>>(nullptr, iters, 12345u);
  else
    fp64_peak_kernel<false><<<sms, warps * 32, 0, st>>>(nullptr, iters, 12345u);
}

static double peak_burst(void* stream, bool fresh) {
  if (stream == nullptr) return -2.0;
  const cudaStream_t st = *static_cast<cudaStream_t*>(stream);
  const int sms = num_sms(), warps = 16, iters = 4096;
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return -30.0;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, st);
    launch_peak(sms, warps, iters, fresh, st);
    cudaEventRecord(e1, st);
    if (cudaEventSynchronize(e1) != cudaSuccess) {
      best = -31.0;
      break;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * 8 * 8 * 4 * 9.0 * iters * (double)warps * sms;
    if (rep > 0 && ms > 0.f) best = std::max(best, flop / (ms * 1e-3) * 1e-9);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return best;
}

// ---- tiled BF16 SpGEMM (smm_bf16_tiled.cuh): driven by the block index instead of parameter stacks -----------------------------
int libsmm_acc_b200_bf16_rk_tile_bytes(int rows) { return ((rows + 7) / 8) * 512; }

int libsmm_acc_b200_bf16_rk_slot_bytes(int rows, int b_operand) { return b_operand ? smm::BT_B_SLOT : ((rows + 7) / 8) * 512; }

int libsmm_acc_b200_pack_bf16_rk(const double* dev_src, int nblks, int rows, int kdim, int row_stride, int k_stride, void* dev_dst,
                                 int dst_pitch, const int* dev_dst_slot, void* stream) {
  if (nblks <= 0) return 0;
  if (stream == nullptr || rows <= 0 || kdim <= 0 || rows > 32 || kdim > 32) return -2;
  if (dst_pitch < ((rows + 7) / 8) * 512 || dst_pitch % 512 != 0) return -2;
  smm::stream_chain_break(*static_cast<cudaStream_t*>(stream));
  const int wpc = 8;
  int grid = (nblks + wpc - 1) / wpc;
  const int max_grid = num_sms() * 8;
  if (grid > max_grid) grid = max_grid;
  smm::pack_bf16_rk_kernel<<<grid, wpc * 32, 0, *static_cast<cudaStream_t*>(stream)>>>(dev_src, nblks, rows, kdim, row_stride, k_stride,
                                                                                    static_cast<unsigned char*>(dev_dst), dst_pitch, dev_dst_slot);
  if (cudaPeekAtLastError() != cudaSuccess) return -31;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int libsmm_acc_b200_bf16_spgemm(const void* a_tiles, const int* dev_a_map, const void* b_tiles, const int* dev_b_map, float* dev_c,
                                const int* dev_c_off, int nrb, int ncb, int nkb, int m, int n, int k, void* stream) {
  if (stream == nullptr || nrb < 0 || ncb < 0 || nkb < 0) return -2;
  if (m <= 0 || n <= 0 || k <= 0 || m > 32 || n > 32 || k > 32) return -10;
  if (nrb == 0 || ncb == 0) return 0;
  const cudaStream_t st = *static_cast<cudaStream_t*>(stream);
  smm::stream_chain_break(st);
  const smm::BtGeom g = smm::bt_geom(m, n);
  const int smem = (int)smm::bt_smem_bytes(g);
  // "bf16_merge" tunable (default 1): adjacent existing B blocks are multiplied by one wide MMA
  // "bf16_a_tmem" tunable: stage the A operand in TMEM (15 block columns per tile instead of 16)
  const int flags = (smm::g_tune.bf16_merge.load(std::memory_order_relaxed) != 0 ? smm::BT_FLAG_MERGE_RUNS : 0) |
                    (smm::g_tune.bf16_a_tmem.load(std::memory_order_relaxed) != 0 ? smm::BT_FLAG_A_TMEM : 0);
  const bool planned = smm::g_tune.bf16_plan.load(std::memory_order_relaxed) != 0;  // the planned kernel has no A-in-TMEM mode
  const int nb = (!planned && (flags & smm::BT_FLAG_A_TMEM)) ? smm::BT_NB_A_TMEM : smm::BT_NB;
  const int bpt = g.bpt;
  const int n_rg = (nrb + bpt - 1) / bpt, n_cg = (ncb + nb - 1) / nb;
  int grid = n_rg * n_cg;
  if (grid > num_sms()) grid = num_sms();
  if (planned) {
    // planned variant: copy commands and MMA runs per (row group | column group, k block) derived once, then the multiply
    static SmemAttrCache smem_set_p;
    int ns = smm::BP_NS;  // DBCSR_B200_BF16_STAGES: experiment knob (fewer pipeline stages)
    if (const char* e = getenv("DBCSR_B200_BF16_STAGES")) ns = std::max(2, std::min(smm::BP_NS, atoi(e)));
    const int smem_p = (int)smm::bp_smem_bytes(ns);
    if (ensure_smem(smm::smm_bf16_planned_kernel, smem_p, smem_set_p) != 0) return -30;
    size_t off[5];
    const size_t bytes = smm::bt_plan_bytes(n_rg, n_cg, nkb, off);
    unsigned char* buf = bt_plan_scratch(st, bytes, off[4]);
    if (buf == nullptr) return -40;
    smm::BtPlanPtrs P;
    P.a_cmd = reinterpret_cast<uint4*>(buf + off[0]);
    P.b_cmd = reinterpret_cast<uint4*>(buf + off[1]);
    P.a_any = buf + off[3];
    P.zeros = buf + off[4];
    const long long items = (long long)(n_rg + n_cg) * nkb;
    int pgrid = (int)std::min<long long>((items + 127) / 128, (long long)num_sms() * 16);
    if (pgrid < 1) pgrid = 1;
    smm::bt_plan_kernel<<<pgrid, 128, 0, st>>>(static_cast<const unsigned char*>(a_tiles), dev_a_map, static_cast<const unsigned char*>(b_tiles),
                                              dev_b_map, nrb, ncb, nkb, m, n, nb, ns, P);
    if (cudaPeekAtLastError() != cudaSuccess) return -31;
    cudaLaunchConfig_t cfgp = {};
    cfgp.gridDim = dim3((unsigned)grid);
    cfgp.blockDim = dim3(smm::BP_THREADS);
    cfgp.dynamicSmemBytes = (size_t)smem_p;
    cfgp.stream = st;
    cudaLaunchAttribute attrp[1];
    attrp[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrp[0].val.programmaticStreamSerializationAllowed = 1;
    cfgp.attrs = attrp;
    cfgp.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfgp, smm::smm_bf16_planned_kernel, P, dev_c, dev_c_off, nrb, ncb, nkb, m, n, ns) != cudaSuccess) return -31;
    g_launches.fetch_add(2, std::memory_order_relaxed);
    return 0;
  }
  static SmemAttrCache smem_set;
  if (ensure_smem(smm::smm_bf16_tiled_kernel, smem, smem_set) != 0) return -30;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(smm::BT_THREADS);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t err = cudaLaunchKernelEx(&cfg, smm::smm_bf16_tiled_kernel, static_cast<const unsigned char*>(a_tiles), dev_a_map,
                                             static_cast<const unsigned char*>(b_tiles), dev_b_map, dev_c, dev_c_off, nrb, ncb, nkb, m, n, flags);
  if (err != cudaSuccess) return -31;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

// Same loop run back to back for `seconds` (power / thermal limits act within tens of milliseconds on a B200: a kernel timed
// inside a long step has to be compared with THIS figure, a kernel timed alone with the burst figure above): throughput over the
// second half of the interval.  Synchronises `stream`.
static double peak_sustained(void* stream, double seconds, bool fresh) {
  if (stream == nullptr || !(seconds > 0.0) || seconds > 10.0) return -2.0;
  const cudaStream_t st = *static_cast<cudaStream_t*>(stream);
  const int sms = num_sms(), warps = 16, iters = 4096;
  const double flop = 2.0 * 8 * 8 * 4 * 9.0 * iters * (double)warps * sms;  // ~1.2 ms per launch at 37 TFLOP/s
  const int n = (int)(seconds / 1.2e-3) + 2;
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return -30.0;
  for (int i = 0; i < n; ++i) {
    if (i == n / 2) cudaEventRecord(e0, st);
    launch_peak(sms, warps, iters, fresh, st);
  }
  cudaEventRecord(e1, st);
  double out = -31.0;
  if (cudaEventSynchronize(e1) == cudaSuccess) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms > 0.f) out = flop * (n - n / 2) / (ms * 1e-3) * 1e-9;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return out;
}

double libsmm_acc_b200_fp64_peak_gflops(void* stream) { return peak_burst(stream, false); }
double libsmm_acc_b200_fp64_peak_sustained_gflops(void* stream, double seconds) { return peak_sustained(stream, seconds, false); }
int libsmm_acc_b200_bf16_tile_bytes(int rows, int kdim) { return ((kdim + 7) / 8) * ((rows + 7) / 8) * 128; }

int libsmm_acc_b200_pack_bf16(const double* dev_src, int nblks, int rows, int kdim, int row_stride, int k_stride, void* dev_dst,
                              void* stream) {
  if (nblks <= 0) return 0;
  if (stream == nullptr || rows <= 0 || kdim <= 0 || rows > 32 || kdim > 32) return -2;
  smm::stream_chain_break(*static_cast<cudaStream_t*>(stream));
  const int wpc = 8;
  int grid = (nblks + wpc - 1) / wpc;
  const int max_grid = num_sms() * 8;
  if (grid > max_grid) grid = max_grid;
  smm::pack_bf16_kernel<<<grid, wpc * 32, 0, *static_cast<cudaStream_t*>(stream)>>>(dev_src, nblks, rows, kdim, row_stride, k_stride,
                                                                                 static_cast<unsigned char*>(dev_dst));
  if (cudaPeekAtLastError() != cudaSuccess) return -31;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int libsmm_acc_b200_kernel_kind(int m, int n, int k, libsmm_acc_data_t datatype) {
  if (datatype == dbcsr_type_bf16_ext) return (m > 0 && n > 0 && k > 0 && m <= 32 && n <= 32 && k <= 32) ? 3 : 0;
  if (m <= 0 || n <= 0 || k <= 0) return 0;
  if (datatype == dbcsr_type_real_4 || datatype == dbcsr_type_complex_4 || datatype == dbcsr_type_complex_8)
    return all_types_enabled() ? 2 : 0;
  if (datatype != dbcsr_type_real_8) return 0;
  return lookup(m, n, k) != nullptr ? 1 : 2;
}

int libsmm_acc_process(const int* host_param_stack, const int* dev_param_stack, int stack_size, libsmm_acc_data_t datatype,
                       const void* dev_a_data, const void* dev_b_data, void* dev_c_data, int m_max, int n_max, int k_max,
                       int max_kernel_dim, c_dbcsr_acc_bool_t def_mnk, void* stack_stream, void* c_stream) {
  if (stack_size < 0 || m_max <= 0 || n_max <= 0 || k_max <= 0) return -2;
  if (stack_stream == nullptr) return -2;
  if (def_mnk != 1) {
    // inhomogeneous stack: the reference returns -1 here (libsmm_acc.cpp:327) and DBCSR falls back to the CPU; this library
    // bins the host stack by shape and drains every bin on the GPU (DBCSR_B200_INHOMOGENEOUS=0 restores the reference behaviour)
    const bool enabled = smm::g_tune.inhomogeneous.load(std::memory_order_relaxed) != 0;
    if (!enabled || datatype != dbcsr_type_real_8) return -1;
    return process_inhomogeneous(host_param_stack, stack_size, static_cast<const double*>(dev_a_data),
                                 static_cast<const double*>(dev_b_data), static_cast<double*>(dev_c_data), max_kernel_dim,
                                 *static_cast<cudaStream_t*>(stack_stream));
  }
  if (datatype == dbcsr_type_bf16_ext) {
    // extension: A/B are BF16 tile panels made by libsmm_acc_b200_pack_bf16, C is FP32; tensor-core (tcgen05) kernel
    if (m_max > 32 || n_max > 32 || k_max > 32) return -10;
    smm::stream_chain_break(*static_cast<cudaStream_t*>(stack_stream));
    const int rc = launch_bf16(dev_param_stack, stack_size, dev_a_data, dev_b_data, dev_c_data, m_max, n_max, k_max,
                               *static_cast<cudaStream_t*>(stack_stream));
    if (rc == 0) g_launches.fetch_add(1, std::memory_order_relaxed);
    return rc;
  }
  if (datatype != dbcsr_type_real_8) {
    // real_4 / complex_4 / complex_8: rejected by the reference (-10, libsmm_acc.cpp:338 => CPU); drained here by the typed
    // generic kernel.  B blocks of these types are transposed by libsmm_acc_transpose below under the same rule as real_8.
    if (!all_types_enabled() || def_mnk != 1) return -10;
    cudaStream_t st = *static_cast<cudaStream_t*>(stack_stream);
    smm::stream_chain_break(st);
    const int bt = (n_max <= max_kernel_dim && k_max <= max_kernel_dim) ? 1 : 0;
    int rc = -10;
    if (datatype == dbcsr_type_real_4)
      rc = launch_generic_typed<float>(dev_param_stack, stack_size, dev_a_data, dev_b_data, dev_c_data, m_max, n_max, k_max, bt, st);
    else if (datatype == dbcsr_type_complex_4)
      rc = launch_generic_typed<float2>(dev_param_stack, stack_size, dev_a_data, dev_b_data, dev_c_data, m_max, n_max, k_max, bt, st);
    else if (datatype == dbcsr_type_complex_8)
      rc = launch_generic_typed<double2>(dev_param_stack, stack_size, dev_a_data, dev_b_data, dev_c_data, m_max, n_max, k_max, bt, st);
    if (rc == 0) {
      g_launches.fetch_add(1, std::memory_order_relaxed);
      return 10;
    }
    return rc;
  }
  const double* a = static_cast<const double*>(dev_a_data);
  const double* b = static_cast<const double*>(dev_b_data);
  double* c = static_cast<double*>(dev_c_data);

  if (m_max > max_kernel_dim || n_max > max_kernel_dim || k_max > max_kernel_dim) {
    // Large blocks.  The reference loops cublasDgemm over the HOST stack on c_stream and synchronises (libsmm_acc.cpp:256-278).
    // Here one launch drains the DEVICE stack asynchronously on stack_stream -- the stream the stack upload was enqueued on and
    // the one DBCSR records the stack buffer's `calculated` event on (src/mm/dbcsr_mm_accdrv.F:512-534), so the buffer and the
    // panels cannot be recycled under the kernel whatever c_stream is; C is accumulated with atomics, so no ordering against
    // c_stream is needed.  B is transposed only if both n and k fit max_kernel_dim (libsmm_acc.cpp:267-270).
    const cudaStream_t big_stream = *static_cast<cudaStream_t*>(stack_stream);
    smm::stream_chain_break(big_stream);
    const int b_transposed = (n_max <= max_kernel_dim && k_max <= max_kernel_dim) ? 1 : 0;
    const int rc = launch_generic(dev_param_stack, stack_size, a, b, c, m_max, n_max, k_max, b_transposed, big_stream);
    if (rc == 0) g_launches.fetch_add(1, std::memory_order_relaxed);
    return rc == 0 ? 10 : rc;
  }

  cudaStream_t stream = *static_cast<cudaStream_t*>(stack_stream);
  const smm::launch_fn fn = lookup(m_max, n_max, k_max);
  if (fn != nullptr) {
    const int rc = fn(dev_param_stack, stack_size, a, b, c, allocation_end(a), allocation_end(b), stream);
    if (rc == 0) g_launches.fetch_add(1, std::memory_order_relaxed);
    return rc;
  }
  smm::stream_chain_break(stream);
  const int rc = rt_eligible(m_max, n_max, k_max)
                   ? launch_rt(dev_param_stack, stack_size, a, b, c, m_max, n_max, k_max, allocation_end(a), allocation_end(b), stream)
                   : big_eligible(m_max, n_max, k_max)
                       ? launch_big(dev_param_stack, stack_size, a, b, c, m_max, n_max, k_max, allocation_end(a), allocation_end(b), stream)
                       : launch_generic(dev_param_stack, stack_size, a, b, c, m_max, n_max, k_max, 1, stream);
  if (rc == 0) g_launches.fetch_add(1, std::memory_order_relaxed);
  return rc == 0 ? 10 : rc;  // 10 = "ran with an untuned kernel" (reference: libsmm_acc.cpp:319)
}

namespace {
int launch_transpose_f64(const int* dev_trs_stack, int stack_size, double* data, int m, int n, const int* dev_trs_blk, float* dev_norms,
                         void* stream) {
  if (stream == nullptr) return -2;
  const size_t blk_bytes = (size_t)m * n * sizeof(double);
  int wpc = (int)((96 * 1024) / blk_bytes);
  if (wpc > 8) wpc = 8;
  if (wpc < 1) return -3;  // cannot happen for m,n <= 80 (51 KB)
  const size_t smem = blk_bytes * wpc;
  static SmemAttrCache attr_set;
  if (ensure_smem(smm::transpose_kernel, (int)smem, attr_set) != 0) return -30;
  int grid = (stack_size + wpc - 1) / wpc;
  const int max_grid = num_sms() * 8;
  if (grid > max_grid) grid = max_grid;
  smm::transpose_kernel<<<grid, wpc * 32, smem, *static_cast<cudaStream_t*>(stream)>>>(dev_trs_stack, stack_size, data, m, n, dev_trs_blk,
                                                                                        dev_norms);
  if (cudaPeekAtLastError() != cudaSuccess) return -31;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}
}  // namespace

int libsmm_acc_transpose(const int* dev_trs_stack, int offset, int stack_size, void* dev_data, libsmm_acc_data_t datatype, int m,
                         int n, int max_kernel_dim, void* stream) {
  if (m > max_kernel_dim || n > max_kernel_dim) return 0;    // reference: libsmm_acc.cpp:485
  if (stack_size <= 0 || m <= 0 || n <= 0) return 0;
  if (stream != nullptr) smm::stream_chain_break(*static_cast<cudaStream_t*>(stream));
  if (datatype != dbcsr_type_real_8) {
    // reference: "transpose not needed" (libsmm_acc.cpp:484) because it never multiplies these types on the device; this
    // library does (typed generic kernel), so their right-panel blocks are transposed exactly like real_8 ones
    if (!all_types_enabled() || stream == nullptr) return 0;
    cudaStream_t st = *static_cast<cudaStream_t*>(stream);
    int rc = 0;
    if (datatype == dbcsr_type_real_4)
      rc = launch_transpose_typed<float>(dev_trs_stack + offset, stack_size, dev_data, m, n, st);
    else if (datatype == dbcsr_type_complex_4)
      rc = launch_transpose_typed<float2>(dev_trs_stack + offset, stack_size, dev_data, m, n, st);
    else if (datatype == dbcsr_type_complex_8)
      rc = launch_transpose_typed<double2>(dev_trs_stack + offset, stack_size, dev_data, m, n, st);
    if (rc == 0) g_launches.fetch_add(1, std::memory_order_relaxed);
    return rc;
  }
  return launch_transpose_f64(dev_trs_stack + offset, stack_size, static_cast<double*>(dev_data), m, n, nullptr, nullptr, stream);
}

// Transpose + block norms in ONE pass over the right panel: like libsmm_acc_transpose for real_8, and in
// addition dev_norms[dev_trs_blk[i]] = sum of squares of block i (float, what c_calculate_norms computes), for every
// i in [offset, offset + stack_size).  Blocks with a dimension above max_kernel_dim are not transposed by DBCSR: -3, nothing done
// (use c_calculate_norms for those).
int libsmm_acc_b200_transpose_norms(const int* dev_trs_stack, const int* dev_trs_blk, int offset, int stack_size, double* dev_data, int m,
                                    int n, int max_kernel_dim, float* dev_norms, void* stream) {
  if (m > max_kernel_dim || n > max_kernel_dim) return -3;
  if (stack_size <= 0 || m <= 0 || n <= 0) return 0;
  if (stream == nullptr || dev_norms == nullptr) return -2;
  smm::stream_chain_break(*static_cast<cudaStream_t*>(stream));
  return launch_transpose_f64(dev_trs_stack + offset, stack_size, dev_data, m, n, dev_trs_blk != nullptr ? dev_trs_blk + offset : nullptr,
                              dev_norms, stream);
}

int c_calculate_norms(const double* mat, int nblks, const int* offsets, const int* nelems, float* norms, void* stream_ptr) {
  if (nblks <= 0) return 0;
  if (stream_ptr == nullptr) return -2;
  smm::stream_chain_break(*static_cast<cudaStream_t*>(stream_ptr));
  const int wpc = 8;
  int grid = (nblks + wpc - 1) / wpc;
  const int max_grid = num_sms() * 8;
  if (grid > max_grid) grid = max_grid;
  smm::norms_kernel<float><<<grid, wpc * 32, 0, *static_cast<cudaStream_t*>(stream_ptr)>>>(mat, nblks, offsets, nelems, norms);
  if (cudaPeekAtLastError() != cudaSuccess) return -31;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int libsmm_acc_b200_block_norms_f64(const double* mat, int nblks, const int* offsets, const int* nelems, double* norms, void* stream_ptr) {
  if (nblks <= 0) return 0;
  if (stream_ptr == nullptr) return -2;
  smm::stream_chain_break(*static_cast<cudaStream_t*>(stream_ptr));
  const int wpc = 8;
  int grid = (nblks + wpc - 1) / wpc;
  const int max_grid = num_sms() * 8;
  if (grid > max_grid) grid = max_grid;
  smm::norms_kernel<double><<<grid, wpc * 32, 0, *static_cast<cudaStream_t*>(stream_ptr)>>>(mat, nblks, offsets, nelems, norms);
  if (cudaPeekAtLastError() != cudaSuccess) return -31;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int libsmm_acc_b200_gather_blocks(const double* src, double* dst, int nblks, const int* src_offsets, const int* dst_offsets,
                                  const int* nelems, void* stream_ptr) {
  if (nblks <= 0) return 0;
  if (stream_ptr == nullptr) return -2;
  smm::stream_chain_break(*static_cast<cudaStream_t*>(stream_ptr));
  const int wpc = 8;
  int grid = (nblks + wpc - 1) / wpc;
  const int max_grid = num_sms() * 8;
  if (grid > max_grid) grid = max_grid;
  smm::gather_blocks_kernel<<<grid, wpc * 32, 0, *static_cast<cudaStream_t*>(stream_ptr)>>>(src, dst, nblks, src_offsets, dst_offsets, nelems);
  if (cudaPeekAtLastError() != cudaSuccess) return -32;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

}  // extern "C"
