// dbcsr_b200/csrc/smm_launch.h -- host-side launcher type shared by the per-shape kernel tables and libsmm_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace smm {

// Enqueue the stack drain on `stream`. Returns 0 or a negative error code; never synchronises.
typedef int (*launch_fn)(const int* dev_stack, int stack_size, const double* a, const double* b, double* c, uint64_t a_limit,
                         uint64_t b_limit, cudaStream_t stream);

// Block sizes with a specialised DMMA kernel per (m,n,k) triplet: the CP2K shapes named by BASELINE.json.
#define SMM_TUNED_SIZES(X) X(5) X(13) X(23) X(26) X(32)

// Programmatic-dependent-launch chain mode (libsmm_acc_b200_stream_chain): true iff the caller declared `stream` a chain of
// independent stack drains AND the previous launch this library put on it was such a drain -- then the kernel may read A/B and
// RED into C without waiting for its predecessor grid.  chain_break(stream) is called by every other launch of the library.
bool stream_chain_mode(cudaStream_t stream);
void stream_chain_break(cudaStream_t stream);

launch_fn lookup_m5(int n, int k);
launch_fn lookup_m13(int n, int k);
launch_fn lookup_m23(int n, int k);
launch_fn lookup_m26(int n, int k);
launch_fn lookup_m32(int n, int k);

}  // namespace smm
