// dbcsr_b200/csrc/acc_runtime.cu -- thin C shim over the CUDA runtime implementing include/dbcsr_acc.h.
//
// Semantics follow the reference's CUDA backend (src/acc/cuda_hip/acc_{init,dev,stream,event,mem,error}.cpp) as pinned by
// its conformance test tests/dbcsr_acc_test.c; the implementation is new.  Differences that matter on B200:
//   * errors are reported through the return code (the reference prints and exit(1)s, src/acc/cuda/acc_cuda.h:29-36);
//     set DBCSR_B200_ABORT_ON_ERROR=1 to get the reference's fail-stop behaviour;
//   * streams get NVTX-free names only (no profiling dependency), priorities are clamped to the device range.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <pthread.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/dbcsr_acc.h"
#include "../../include/dbcsr_acc_libsmm.h"

namespace {

std::atomic<int> g_init_count{0};

bool abort_on_error() {
  static const bool v = [] {
    const char* e = getenv("DBCSR_B200_ABORT_ON_ERROR");
    return e != nullptr && atoi(e) != 0;
  }();
  return v;
}

int check(cudaError_t err, const char* what) {
  if (err == cudaSuccess) return 0;
  fprintf(stderr, "dbcsr_acc_b200: %s failed: %s\n", what, cudaGetErrorString(err));
  if (abort_on_error()) exit(1);
  return -1;
}

#define ACC_TRY(call) \
  do { \
    if (check((call), #call) != 0) return -1; \
  } while (0)

inline cudaStream_t as_stream(void* h) { return *static_cast<cudaStream_t*>(h); }
inline cudaEvent_t as_event(void* h) { return *static_cast<cudaEvent_t*>(h); }

}  // namespace

extern "C" {

// Weak fall-backs for the timer call-backs DBCSR normally provides (acc.h:73-74).
__attribute__((weak)) void c_dbcsr_timeset(const char** routineN, const int* routineN_len, int* handle) {
  (void)routineN;
  (void)routineN_len;
  if (handle != nullptr) *handle = 0;
}
__attribute__((weak)) void c_dbcsr_timestop(const int* handle) { (void)handle; }

int c_dbcsr_acc_init(void) {
  // Establish the primary context of the active device (reference: cuInit + cuDevicePrimaryCtxRetain).
  ACC_TRY(cudaFree(nullptr));
  g_init_count.fetch_add(1);
  return libsmm_acc_init();
}

int c_dbcsr_acc_finalize(void) {
  if (g_init_count.load() > 0) g_init_count.fetch_sub(1);
  return libsmm_acc_finalize();
}

void c_dbcsr_acc_clear_errors(void) { (void)cudaGetLastError(); }

int c_dbcsr_acc_get_ndevices(int* ndevices) {
  if (ndevices == nullptr) return -1;
  int n = 0;
  const cudaError_t err = cudaGetDeviceCount(&n);
  if (err == cudaErrorNoDevice || err == cudaErrorInsufficientDriver) {
    // a machine without GPUs is not an error for the caller (tests/dbcsr_acc_test.c:86 continues with 0 devices)
    (void)cudaGetLastError();
    *ndevices = 0;
    return 0;
  }
  ACC_TRY(err);
  *ndevices = n;
  return 0;
}

int c_dbcsr_acc_set_active_device(int device_id) {
  int current = -1;
  ACC_TRY(cudaSetDevice(device_id));
  ACC_TRY(cudaGetDevice(&current));
  if (current != device_id) return -1;
  ACC_TRY(cudaFree(nullptr));  // establish the context now, like the reference (acc_dev.cpp:44)
  return 0;
}

int c_dbcsr_acc_device_synchronize(void) {
  ACC_TRY(cudaDeviceSynchronize());
  return 0;
}

int c_dbcsr_acc_stream_priority_range(int* least, int* greatest) {
  int lo = -1, hi = -1;
  ACC_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  if (least != nullptr) *least = lo;
  if (greatest != nullptr) *greatest = hi;
  return 0;
}

int c_dbcsr_acc_stream_create(void** stream_p, const char* name, int priority) {
  (void)name;
  if (stream_p == nullptr) return -1;
  cudaStream_t* s = static_cast<cudaStream_t*>(malloc(sizeof(cudaStream_t)));
  if (s == nullptr) return -1;
  cudaError_t err;
  if (priority > 0) {
    err = cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, priority);
  }
  else {
    err = cudaStreamCreate(s);
  }
  if (check(err, "cudaStreamCreate") != 0) {
    free(s);
    *stream_p = nullptr;
    return -1;
  }
  *stream_p = s;
  return 0;
}

int c_dbcsr_acc_stream_destroy(void* stream) {
  c_dbcsr_acc_clear_errors();
  if (stream == nullptr) return 0;
  const cudaError_t err = cudaStreamDestroy(as_stream(stream));
  free(stream);
  return check(err, "cudaStreamDestroy");
}

int c_dbcsr_acc_stream_sync(void* stream) {
  c_dbcsr_acc_clear_errors();
  if (stream == nullptr) return -1;
  ACC_TRY(cudaStreamSynchronize(as_stream(stream)));
  return 0;
}

int c_dbcsr_acc_stream_wait_event(void* stream, void* event) {
  if (stream == nullptr || event == nullptr) return -1;
  ACC_TRY(cudaStreamWaitEvent(as_stream(stream), as_event(event), 0));
  return 0;
}

int c_dbcsr_acc_event_create(void** event_p) {
  if (event_p == nullptr) return -1;
  cudaEvent_t* e = static_cast<cudaEvent_t*>(malloc(sizeof(cudaEvent_t)));
  if (e == nullptr) return -1;
  // timing disabled: these events only order work (cheaper record/query); DBCSR never asks for elapsed time
  if (check(cudaEventCreateWithFlags(e, cudaEventDisableTiming), "cudaEventCreate") != 0) {
    free(e);
    *event_p = nullptr;
    return -1;
  }
  *event_p = e;
  return 0;
}

int c_dbcsr_acc_event_destroy(void* event) {
  c_dbcsr_acc_clear_errors();
  if (event == nullptr) return 0;
  const cudaError_t err = cudaEventDestroy(as_event(event));
  free(event);
  return check(err, "cudaEventDestroy");
}

int c_dbcsr_acc_event_record(void* event, void* stream) {
  if (event == nullptr || stream == nullptr) return -1;
  ACC_TRY(cudaEventRecord(as_event(event), as_stream(stream)));
  return 0;
}

int c_dbcsr_acc_event_query(void* event, c_dbcsr_acc_bool_t* has_occurred) {
  if (event == nullptr || has_occurred == nullptr) return -1;
  const cudaError_t err = cudaEventQuery(as_event(event));
  if (err == cudaSuccess) {
    *has_occurred = 1;
    return 0;
  }
  if (err == cudaErrorNotReady) {
    (void)cudaGetLastError();
    *has_occurred = 0;
    return 0;
  }
  return check(err, "cudaEventQuery");
}

int c_dbcsr_acc_event_synchronize(void* event) {
  if (event == nullptr) return -1;
  ACC_TRY(cudaEventSynchronize(as_event(event)));
  return 0;
}

int c_dbcsr_acc_dev_mem_allocate(void** dev_mem, size_t nbytes) {
  if (dev_mem == nullptr) return -2;
  ACC_TRY(cudaMalloc(dev_mem, nbytes));
  return 0;
}

int c_dbcsr_acc_dev_mem_deallocate(void* dev_mem) {
  ACC_TRY(cudaFree(dev_mem));
  return 0;
}

int c_dbcsr_acc_dev_mem_set_ptr(void** dev_mem, void* other, size_t lb) {
  if (dev_mem == nullptr) return -1;
  *dev_mem = static_cast<char*>(other) + lb;
  return 0;
}

int c_dbcsr_acc_host_mem_allocate(void** host_mem, size_t nbytes, void* stream) {
  (void)stream;
  if (host_mem == nullptr) return -2;
  ACC_TRY(cudaHostAlloc(host_mem, nbytes, cudaHostAllocDefault));
  return 0;
}

int c_dbcsr_acc_host_mem_deallocate(void* host_mem, void* stream) {
  (void)stream;
  ACC_TRY(cudaFreeHost(host_mem));
  return 0;
}

int c_dbcsr_acc_memcpy_h2d(const void* host_mem, void* dev_mem, size_t nbytes, void* stream) {
  if (stream == nullptr) return -1;
  if (nbytes == 0) return 0;  // empty panels / stacks: nothing to enqueue
  ACC_TRY(cudaMemcpyAsync(dev_mem, host_mem, nbytes, cudaMemcpyHostToDevice, as_stream(stream)));
  return 0;
}

int c_dbcsr_acc_memcpy_d2h(const void* dev_mem, void* host_mem, size_t nbytes, void* stream) {
  if (stream == nullptr) return -1;
  if (nbytes == 0) return 0;
  ACC_TRY(cudaMemcpyAsync(host_mem, dev_mem, nbytes, cudaMemcpyDeviceToHost, as_stream(stream)));
  return 0;
}

int c_dbcsr_acc_memcpy_d2d(const void* devmem_src, void* devmem_dst, size_t nbytes, void* stream) {
  if (stream == nullptr) {
    ACC_TRY(cudaMemcpy(devmem_dst, devmem_src, nbytes, cudaMemcpyDeviceToDevice));
  }
  else {
    ACC_TRY(cudaMemcpyAsync(devmem_dst, devmem_src, nbytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
  }
  return 0;
}

int c_dbcsr_acc_memset_zero(void* dev_mem, size_t offset, size_t nbytes, void* stream) {
  void* p = static_cast<char*>(dev_mem) + offset;
  if (stream == nullptr) {
    ACC_TRY(cudaMemset(p, 0, nbytes));
  }
  else {
    ACC_TRY(cudaMemsetAsync(p, 0, nbytes, as_stream(stream)));
  }
  return 0;
}

int c_dbcsr_acc_dev_mem_info(size_t* mem_free, size_t* mem_total) {
  size_t f = 0, t = 0;
  ACC_TRY(cudaMemGetInfo(&f, &t));
  if (mem_free != nullptr) *mem_free = f;
  if (mem_total != nullptr) *mem_total = t;
  return 0;
}

// ---- optional profiling hooks (src/acc/cuda/dbcsr_cuda_profiling.F:31-56 binds them when DBCSR is built with __CUDA_PROFILING):
// NVTX3 is header-only and resolves the tool's injection library at run time, so there is no link dependency and the calls are
// no-ops without a profiler attached.  The colour is a stable function of the message, the payload carries its length.
int cuda_nvtx_range_push_cu(const char* message) {
  static const uint32_t palette[8] = {0xFF4E79A7, 0xFFF28E2B, 0xFFE15759, 0xFF76B7B2, 0xFF59A14F, 0xFFEDC948, 0xFFB07AA1, 0xFF9C755F};
  if (message == nullptr) message = "";
  uint32_t h = 2166136261u;
  size_t len = 0;
  for (const char* c = message; *c != 0; ++c, ++len) h = (h ^ (uint32_t)(unsigned char)*c) * 16777619u;
  nvtxEventAttributes_t a;
  memset(&a, 0, sizeof(a));
  a.version = NVTX_VERSION;
  a.size = NVTX_EVENT_ATTRIB_STRUCT_SIZE;
  a.colorType = NVTX_COLOR_ARGB;
  a.color = palette[h & 7u];
  a.messageType = NVTX_MESSAGE_TYPE_ASCII;
  a.message.ascii = message;
  a.payloadType = NVTX_PAYLOAD_TYPE_UNSIGNED_INT64;
  a.payload.ullValue = (uint64_t)len;
  return nvtxRangePushEx(&a);
}
int cuda_nvtx_range_pop_cu(void) { return nvtxRangePop(); }
void cuda_nvtx_name_osthread_cu(char* name) {
  if (name != nullptr) nvtxNameOsThreadA((uint32_t)(uintptr_t)pthread_self(), name);
}

}  // extern "C"
