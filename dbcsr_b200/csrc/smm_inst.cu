// dbcsr_b200/csrc/smm_inst.cu -- instantiates smm_dmma_kernel<SMM_M, n, k> for every (n,k) of the tuned size set and
// exports smm::lookup_m<SMM_M>(n,k).  Compiled once per SMM_M (see Makefile) so the 125 kernels build in parallel.
// This is the ahead-of-time replacement of the reference's NVRTC JIT + kernel cache (src/acc/libsmm_acc/libsmm_acc.cpp:90-253).
#include <atomic>
#include <cstdio>
#include <cstdlib>

#include "smm_dmma.cuh"
#include "smm_launch.h"

#ifndef SMM_M
#  error "compile with -DSMM_M=<block rows>"
#endif

namespace smm {
namespace {

int g_num_sms = 0;
// DBCSR_B200_PDL=0 disables programmatic dependent launch (A/B experiments); default on
const bool g_use_pdl = [] {
  const char* e = getenv("DBCSR_B200_PDL");
  return e == nullptr || atoi(e) != 0;
}();

template <int M, int N, int K>
int launch(const int* dev_stack, int stack_size, const double* a, const double* b, double* c, uint64_t a_limit, uint64_t b_limit,
           cudaStream_t stream) {
  using SH = Shape<M, N, K>;
#if defined(SMM_FORCE_NST) && defined(SMM_FORCE_WPC)
  constexpr int NST = (SMM_FORCE_WPC * SMM_FORCE_NST * SH::STAGE <= 220 * 1024) ? SMM_FORCE_NST : pick_nst(SH::STAGE);
  constexpr int WPC = (SMM_FORCE_WPC * SMM_FORCE_NST * SH::STAGE <= 220 * 1024) ? SMM_FORCE_WPC : pick_wpc(SH::STAGE);
#else
  constexpr int NST = pick_nst(SH::STAGE);
  constexpr int WPC = pick_wpc(SH::STAGE);
#endif
  constexpr int SMEM = round_up_c(WPC * NST * 8, 128) + WPC * NST * SH::STAGE;
  static_assert(SMEM <= 227 * 1024, "shared memory budget exceeded");
  auto kern = smm_dmma_kernel<M, N, K, NST, WPC>;
  static std::atomic<int> ctas_per_sm{0};
  int cps = ctas_per_sm.load(std::memory_order_acquire);
  if (cps == 0) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) return -30;
    int dev = 0, nb = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -30;
    if (g_num_sms == 0) cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, WPC * 32, SMEM) != cudaSuccess || nb < 1) return -30;
    cps = nb;
    ctas_per_sm.store(cps, std::memory_order_release);
  }
  if (stack_size <= 0) return 0;
  const int max_grid = g_num_sms * cps;
  // at least 4 entries per warp so that the pipeline prologue and the C flush are amortised
  int grid = (stack_size + WPC * 4 - 1) / (WPC * 4);
  if (grid > max_grid) grid = max_grid;
  const int warps = grid * WPC;
  const int chunk = (stack_size + warps - 1) / warps;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(WPC * 32);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const unsigned long long al = a_limit, bl = b_limit;
  const cudaError_t err = cudaLaunchKernelEx(&cfg, kern, dev_stack, stack_size, a, b, c, al, bl, chunk);
  return (err == cudaSuccess) ? 0 : -31;
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
