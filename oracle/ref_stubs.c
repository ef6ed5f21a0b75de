/*
 * oracle/ref_stubs.c -- TEST INFRASTRUCTURE.  The reference's libsmm_acc_benchmark.cpp also contains a GPU
 * benchmark driver that calls the CUDA driver/runtime API.  oracle/_ref only uses its pure-CPU checker
 * functions, and must load on machines without a CUDA driver (ctypes binds RTLD_NOW), so the CUDA entry
 * points that file references are satisfied by aborting stubs.  None of them is reachable from ref_shim.cpp.
 */
#include <stdio.h>
#include <stdlib.h>
#define STUB(name) \
  int name(void) { \
    fprintf(stderr, "oracle/_ref: %s is a stub (CPU checker only)\n", #name); \
    abort(); \
    return -1; \
  }
STUB(cuGetErrorName)
STUB(cuEventCreate)
STUB(cuEventDestroy_v2)
STUB(cuEventElapsedTime)
STUB(cuEventRecord)
STUB(cuEventSynchronize)
STUB(cuStreamCreate)
STUB(cuLaunchKernel)
STUB(cudaFree)
STUB(cudaGetErrorName)
STUB(cudaMalloc)
STUB(cudaMemcpy)
STUB(cudaMemset)
