// dbcsr_b200/csrc/host/replay.cu -- one whole Cannon multiply of a rank on PRE-BUILT device stacks, enqueued by one C call
// (include/dbcsr_b200_host.h, dbcsr_b200_replay_*): the multi-GPU counterpart of "stack-kernel only" timing
// (src/acc/acc_bench.c:338-345) -- per tick the peer pulls of the panels (cudaMemcpyAsync over NVLink: copy engines, no SM) on a
// side stream, ordered by events against the stack kernels of the tick on the compute stream, C zeroed once per multiply.
// Issuing the ~50 operations of a step from C instead of a Python/ctypes loop takes the host out of the critical path at 8 GPUs,
// where a whole step lasts ~1.5 ms (profiles/r01_cannon_trace_n8_p2p.txt).  Schedule = dbcsr_b200/cannon.py (Cannon with virtual
// k-slices, src/mm/dbcsr_mm_cannon.F:839-1771); this file only replays it.
#include <cuda_runtime.h>

#include <vector>

#include "../../../include/dbcsr_acc.h"
#include "../../../include/dbcsr_acc_libsmm.h"
#include "../../../include/dbcsr_b200_host.h"

namespace {
struct Pull {
  void* dst;
  const void* src;
  size_t bytes;
};
struct Stack {
  const int* dev;
  int size, m, n, k, defined;
};
struct Tick {
  std::vector<Pull> pulls;
  std::vector<Stack> stacks;
  const void *a = nullptr, *b = nullptr;
  cudaEvent_t pulled = nullptr, computed = nullptr;
  bool computed_once = false;
};
}  // namespace

struct dbcsr_b200_replay {
  std::vector<Tick> ticks;  // in the order the rank takes them
  void* c[2] = {nullptr, nullptr};
  size_t c_bytes = 0;
  int zero_overlap = 1;
  cudaStream_t pull_stream = nullptr, zero_stream = nullptr;
  void *pull_handle = nullptr, *zero_handle = nullptr;  // acc-style handles (pointer to the stream) of the two side streams
  cudaEvent_t zeroed[2] = {nullptr, nullptr}, freed[2] = {nullptr, nullptr};
  bool freed_once[2] = {false, false}, zeroed_once[2] = {false, false};
  long long step_no = 0;
  int current = 0;
};

extern "C" {

dbcsr_b200_replay_t* dbcsr_b200_replay_create(int nticks) {
  if (nticks < 1) return nullptr;
  auto* r = new dbcsr_b200_replay();
  r->ticks.resize((size_t)nticks);
  bool ok = cudaStreamCreateWithFlags(&r->pull_stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&r->zero_stream, cudaStreamNonBlocking) == cudaSuccess;
  for (auto& t : r->ticks)
    ok = ok && cudaEventCreateWithFlags(&t.pulled, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&t.computed, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; i < 2; ++i)
    ok = ok && cudaEventCreateWithFlags(&r->zeroed[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&r->freed[i], cudaEventDisableTiming) == cudaSuccess;
  r->pull_handle = &r->pull_stream;
  r->zero_handle = &r->zero_stream;
  if (!ok) {
    dbcsr_b200_replay_destroy(r);
    return nullptr;
  }
  return r;
}

void dbcsr_b200_replay_destroy(dbcsr_b200_replay_t* r) {
  if (r == nullptr) return;
  if (r->pull_stream != nullptr) cudaStreamSynchronize(r->pull_stream);
  if (r->zero_stream != nullptr) cudaStreamSynchronize(r->zero_stream);
  for (auto& t : r->ticks) {
    if (t.pulled != nullptr) cudaEventDestroy(t.pulled);
    if (t.computed != nullptr) cudaEventDestroy(t.computed);
  }
  for (int i = 0; i < 2; ++i) {
    if (r->zeroed[i] != nullptr) cudaEventDestroy(r->zeroed[i]);
    if (r->freed[i] != nullptr) cudaEventDestroy(r->freed[i]);
  }
  if (r->pull_stream != nullptr) cudaStreamDestroy(r->pull_stream);
  if (r->zero_stream != nullptr) cudaStreamDestroy(r->zero_stream);
  delete r;
}

int dbcsr_b200_replay_set_panels(dbcsr_b200_replay_t* r, int tick, const void* a_dev, const void* b_dev) {
  if (r == nullptr || tick < 0 || tick >= (int)r->ticks.size()) return -1;
  r->ticks[(size_t)tick].a = a_dev;
  r->ticks[(size_t)tick].b = b_dev;
  return 0;
}

int dbcsr_b200_replay_add_pull(dbcsr_b200_replay_t* r, int tick, void* dst, const void* src, size_t bytes) {
  if (r == nullptr || tick < 0 || tick >= (int)r->ticks.size() || dst == nullptr || src == nullptr) return -1;
  if (bytes > 0) r->ticks[(size_t)tick].pulls.push_back({dst, src, bytes});
  return 0;
}

int dbcsr_b200_replay_add_stack(dbcsr_b200_replay_t* r, int tick, const int* dev_stack, int size, int m, int n, int k, int defined_mnk) {
  if (r == nullptr || tick < 0 || tick >= (int)r->ticks.size() || size < 0) return -1;
  if (size > 0) r->ticks[(size_t)tick].stacks.push_back({dev_stack, size, m, n, k, defined_mnk});
  return 0;
}

int dbcsr_b200_replay_set_c(dbcsr_b200_replay_t* r, void* c0, void* c1, size_t bytes, int zero_overlap) {
  if (r == nullptr || c0 == nullptr) return -1;
  r->c[0] = c0;
  r->c[1] = c1 != nullptr ? c1 : c0;
  r->c_bytes = bytes;
  r->zero_overlap = (zero_overlap != 0 && c1 != nullptr && c1 != c0) ? 1 : 0;
  return 0;
}

void* dbcsr_b200_replay_current_c(const dbcsr_b200_replay_t* r) { return r == nullptr ? nullptr : r->c[r->current]; }

int dbcsr_b200_replay_step(dbcsr_b200_replay_t* r, void* compute_stream) {
  if (r == nullptr || compute_stream == nullptr || r->c[0] == nullptr) return -1;
  const cudaStream_t cs = *static_cast<cudaStream_t*>(compute_stream);
  const int k = (int)(r->step_no & 1);
  ++r->step_no;
  r->current = r->zero_overlap ? k : 0;
  void* c = r->c[r->current];
  if (r->zero_overlap) {
    // the buffer of the NEXT step is zeroed on a side stream while this step's stacks run; the step owns the memset it issued
    if (r->freed_once[1 - k] && cudaStreamWaitEvent(r->zero_stream, r->freed[1 - k], 0) != cudaSuccess) return -2;
    if (c_dbcsr_acc_memset_zero(r->c[1 - k], 0, r->c_bytes, r->zero_handle) != 0) return -2;
    if (cudaEventRecord(r->zeroed[1 - k], r->zero_stream) != cudaSuccess) return -2;
    r->zeroed_once[1 - k] = true;
    if (r->zeroed_once[k]) {
      if (cudaStreamWaitEvent(cs, r->zeroed[k], 0) != cudaSuccess) return -2;
    }
    else if (c_dbcsr_acc_memset_zero(c, 0, r->c_bytes, compute_stream) != 0) {  // very first step: nobody zeroed this buffer yet
      return -2;
    }
  }
  else if (c_dbcsr_acc_memset_zero(c, 0, r->c_bytes, compute_stream) != 0) {
    return -2;
  }
  auto post_pulls = [&](Tick& t) -> int {
    // the receive buffers of this tick were last read by the previous multiply's kernels of the same tick
    if (t.computed_once && cudaStreamWaitEvent(r->pull_stream, t.computed, 0) != cudaSuccess) return -3;
    for (const Pull& p : t.pulls)
      if (cudaMemcpyAsync(p.dst, p.src, p.bytes, cudaMemcpyDefault, r->pull_stream) != cudaSuccess) return -3;
    return cudaEventRecord(t.pulled, r->pull_stream) == cudaSuccess ? 0 : -3;
  };
  const size_t nt = r->ticks.size();
  // every pull of this multiply is posted up front (one receive buffer per tick): a pull starts as soon as the previous multiply's
  // kernels of the same tick have released the buffer, i.e. beside the LAST ticks of the previous multiply when steps follow each other
  for (size_t u = 0; u < nt; ++u)
    if (post_pulls(r->ticks[u]) != 0) return -3;
  for (size_t it = 0; it < nt; ++it) {
    Tick& t = r->ticks[it];
    if (cudaStreamWaitEvent(cs, t.pulled, 0) != cudaSuccess) return -4;
    for (const Stack& s : t.stacks) {
      const int rc = libsmm_acc_process(nullptr, s.dev, s.size, dbcsr_type_real_8, t.a, t.b, c, s.m, s.n, s.k, 80, s.defined, compute_stream,
                                        compute_stream);
      if (rc < 0) return rc;
    }
    if (cudaEventRecord(t.computed, cs) != cudaSuccess) return -4;
    t.computed_once = true;
  }
  if (r->zero_overlap) {
    if (cudaEventRecord(r->freed[k], cs) != cudaSuccess) return -5;
    r->freed_once[k] = true;
    if (cudaStreamWaitEvent(cs, r->zeroed[1 - k], 0) != cudaSuccess) return -5;
  }
  return 0;
}

}  // extern "C"
