/*
 * include/dbcsr_acc.h -- C ABI of the B200-native DBCSR accelerator runtime (libdbcsr_acc_b200.so).
 *
 * Drop-in boundary, part 1 of 2: the 26 entry points DBCSR's Fortran layer binds through ISO_C_BINDING
 * (reference interface: src/acc/acc.h:34-71; callers src/acc/dbcsr_acc_{init,device,stream,event,devmem,hostmem}.F).
 * Every function returns 0 (EXIT_SUCCESS) on success unless stated otherwise.  Handles are opaque `void*`:
 *   stream handle = pointer to a heap-allocated cudaStream_t   (reference: src/acc/cuda_hip/acc_stream.cpp:40-42)
 *   event  handle = pointer to a heap-allocated cudaEvent_t    (reference: src/acc/cuda_hip/acc_event.cpp:25-26)
 * so code that dereferences a handle as `*(cudaStream_t*)h` keeps working.  No torch / C++ types cross this ABI.
 */
#ifndef DBCSR_B200_ACC_H
#define DBCSR_B200_ACC_H

#include <stddef.h>

#if defined(__cplusplus)
extern "C" {
#endif

typedef int c_dbcsr_acc_bool_t; /* acc.h:31 */

/* -- initialisation / finalisation (acc.h:34-35; src/acc/cuda_hip/acc_init.cpp:22-50) -------------------------
 * init retains the primary context of the active device and calls libsmm_acc_init(); both are idempotent. */
int c_dbcsr_acc_init(void);
int c_dbcsr_acc_finalize(void);

/* -- error handling (acc.h:38; src/acc/cuda_hip/acc_error.cpp): swallow a pending sticky-free CUDA error. */
void c_dbcsr_acc_clear_errors(void);

/* -- devices (acc.h:41-43; src/acc/cuda_hip/acc_dev.cpp:24-55). get_ndevices / set_active_device are legal BEFORE
 * init (tests/dbcsr_acc_test.c:77-82). set_active_device returns -1 if the device could not be made current. */
int c_dbcsr_acc_get_ndevices(int* ndevices);
int c_dbcsr_acc_set_active_device(int device_id);
int c_dbcsr_acc_device_synchronize(void);

/* -- streams (acc.h:46-52; src/acc/cuda_hip/acc_stream.cpp:30-104). priority > 0 => non-blocking stream with that
 * priority, else a default (blocking) stream; name may be NULL or ""; destroy(NULL) is not an error. */
int c_dbcsr_acc_stream_priority_range(int* least, int* greatest);
int c_dbcsr_acc_stream_create(void** stream_p, const char* name, int priority);
int c_dbcsr_acc_stream_destroy(void* stream);
int c_dbcsr_acc_stream_sync(void* stream);
int c_dbcsr_acc_stream_wait_event(void* stream, void* event);

/* -- events (acc.h:55-59; src/acc/cuda_hip/acc_event.cpp:25-104). An unrecorded event queries as occurred
 * (tests/dbcsr_acc_test.c:138-142); create/destroy are thread-safe; destroy(NULL) is not an error. */
int c_dbcsr_acc_event_create(void** event_p);
int c_dbcsr_acc_event_destroy(void* event);
int c_dbcsr_acc_event_record(void* event, void* stream);
int c_dbcsr_acc_event_query(void* event, c_dbcsr_acc_bool_t* has_occurred);
int c_dbcsr_acc_event_synchronize(void* event);

/* -- memory (acc.h:62-71; src/acc/cuda_hip/acc_mem.cpp:29-142). Device pointers are raw and owned by the caller;
 * host memory is pinned; copies are asynchronous on *stream; d2d and memset_zero accept stream == NULL (synchronous). */
int c_dbcsr_acc_dev_mem_allocate(void** dev_mem, size_t nbytes);
int c_dbcsr_acc_dev_mem_deallocate(void* dev_mem);
int c_dbcsr_acc_dev_mem_set_ptr(void** dev_mem, void* other, size_t lb);
int c_dbcsr_acc_host_mem_allocate(void** host_mem, size_t nbytes, void* stream);
int c_dbcsr_acc_host_mem_deallocate(void* host_mem, void* stream);
int c_dbcsr_acc_memcpy_h2d(const void* host_mem, void* dev_mem, size_t nbytes, void* stream);
int c_dbcsr_acc_memcpy_d2h(const void* dev_mem, void* host_mem, size_t nbytes, void* stream);
int c_dbcsr_acc_memcpy_d2d(const void* devmem_src, void* devmem_dst, size_t nbytes, void* stream);
int c_dbcsr_acc_memset_zero(void* dev_mem, size_t offset, size_t nbytes, void* stream);
int c_dbcsr_acc_dev_mem_info(size_t* mem_free, size_t* mem_total);

/* -- timer call-backs (acc.h:73-74): IMPORTED by the backend, implemented by DBCSR (src/acc/dbcsr_acc_timings.F:23,44).
 * The library carries weak no-op definitions so that it also links stand-alone (cf. src/acc/libsmm_acc/libsmm_acc_init.cpp:22-35). */
void c_dbcsr_timeset(const char** routineN, const int* routineN_len, int* handle);
void c_dbcsr_timestop(const int* handle);

/* -- optional profiling hooks, bound by DBCSR when it is built with __CUDA_PROFILING (src/acc/cuda/dbcsr_cuda_profiling.F:31-56,
 * reference implementation src/acc/cuda/dbcsr_cuda_nvtx_cu.cpp): NVTX ranges around DBCSR's timed routines and a name for the
 * calling OS thread.  push/pop return the nesting level; all three are no-ops when no profiler is attached. */
int cuda_nvtx_range_push_cu(const char* message);
int cuda_nvtx_range_pop_cu(void);
void cuda_nvtx_name_osthread_cu(char* name);

#if defined(__cplusplus)
}
#endif

#endif /* DBCSR_B200_ACC_H */
