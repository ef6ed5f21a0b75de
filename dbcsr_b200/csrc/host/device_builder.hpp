// dbcsr_b200/csrc/host/device_builder.hpp -- device-side stack builder.
//
// What the host path does one product at a time (csr_multiply_low: src/mm/dbcsr_mm_csr.F:178-359; hash tables :294-323; stack
// flush :704-739; stack_sort src/mm/dbcsr_mm_accdrv.F:364-384) is restated as data-parallel passes over the products of one
// Cannon tick, with results IDENTICAL to LocalMultiply (stack_builder.cpp): the same C blocks in the same first-touch order at the
// same offsets, the same stacks with the same entries dispatched in the same order, each in the same device order.
//
//   host   : rec_sort_index of both lists and the recursion of sparse_multrec down to its leaves (cuts only; ~1e3 leaves)
//   device : per leaf range a stable sort by row (the leaf's CSR index), products per (leaf, A block) counted by binary search,
//            exclusive scan = position of every product in the reference's traversal order; C blocks by a table keyed
//            (row, col) whose value is the MINIMUM traversal position touching it (atomicMin) => first-touch flags => scans give
//            the block number and the offset; products partitioned by stack number (stable radix sort); the "stack full ->
//            flush everything above 3/4" rule replayed on the per-stack position lists; every dispatched stack sorted by c_first
//            (one stable radix sort over (stack, c_first) keys) and written as the 3-wide device stack in place on the device.
//   The host receives two small messages per tick (product count; dispatch list + sizes) and the new part of the C index.
//
// The passes are element-wise functors plus scans and sorts; `DeviceExec` runs them as CUDA kernels (cub for scan / radix sort),
// `HostExec` runs the SAME functors in plain loops.  HostExec exists for the CPU test-suite only (the container that builds this
// library has no GPU): the engine never selects it.
#pragma once
#include <cstdint>
#include <utility>
#include <vector>

#include "stack_builder.hpp"

namespace dbcsr_b200 {

struct DevDispatch {
  int ws = 0;              // stack number (1-based, LocalMultiply::descr)
  int begin = 0;           // rank of the stack's first entry among the products of this stack number (this tick)
  int size = 0;            // entries
  int slice = 0;           // row slice (multiply call) the stack was dispatched in
  long long seq_start = 0; // first entry in the tick's device stack array (entries, not ints)
};

struct DevBuildResult {
  std::vector<DevDispatch> dispatch;     // in dispatch order
  std::vector<int> slice_datasize;       // datasize after every slice
  long long nprod = 0;                   // products of this tick
  int nblk_before = 0, nblk_after = 0;
  int datasize_before = 0, datasize_after = 0;
};

// What csr_multiply_low tests per product before it looks the C block up (src/mm/dbcsr_mm_csr.F:270-292, :307), for one tick
struct DevBuildOptions {
  // on-the-fly filter: norms aligned with the SORTED lists, thresholds per local block row; all three or none
  const float* a_norms = nullptr;
  const float* b_norms = nullptr;
  const float* row_eps = nullptr;
  // product with symmetry: skip (row, col) when global row != global col and checker_tr(global row, global col)
  bool c_sym = false;
  const int* global_rows = nullptr;  // per local block row (nullptr: identity)
  const int* global_cols = nullptr;
  // retain_sparsity: products whose C block does not exist are dropped, no block is created
  bool keep_sparsity = false;
};

class IDeviceBuilder {
 public:
  virtual ~IDeviceBuilder() {}
  // forget the product index (new multiply); keeps every allocation
  virtual int reset() = 0;
  // tile > 0: NOT the reference's stacks -- inside every (slice, stack number) group the products are ordered by tile x tile squares
  // of C blocks and by c_first inside a square, then cut into stacks of mm_stack_size entries.  The C index (block order, offsets)
  // and the set of products stay those of the reference; what changes is the order of summation (all products of a C block are
  // adjacent: one accumulation run per block) and the memory locality (a square's A rows, B columns and C blocks fit into L2).
  // Applies when every dispatched stack is of the kind the accelerator driver sorts by c_first; 0 = reference order.
  virtual void set_tile_order(int tile) = 0;
  // One tick of one host thread.  a_sorted / b_sorted: rec-sorted lists; slices: the (a_first, a_last) ranges (1-based, inclusive)
  // this thread multiplies, in order - each is one LocalMultiply::multiply call, i.e. ends with a purge.  mm supplies the
  // recursion (plan) and the block-size / stack maps.  Returns 0 or a negative code.
  virtual int build(LocalMultiply& mm, const Idx3* a_sorted, int na, const std::vector<std::pair<int, int>>& slices, const Idx3* b_sorted,
                    int nb, DevBuildResult& out, const DevBuildOptions& opt = DevBuildOptions()) = 0;
  // Existing C blocks (beta != 0 / retain_sparsity): the work index starts from them (fill_hash_tables, src/mm/dbcsr_mm_csr.F:540-576);
  // after reset() and before the first build.  blk_p: their 1-based offsets; datasize: elements they occupy.
  virtual int preset(int nrows, int ncols, const int* rows, const int* cols, const int* blk_p, int nblks, int datasize) = 0;
  // 3-wide stack of a dispatch entry, in device order where the stack is sorted by c_first on the device (`device_ordered`), else
  // in traversal order (the caller orders it on the host: binning, inhomogeneous stacks)
  virtual const int* stack3(const DevDispatch& d) const = 0;       // address in the executor's memory space
  virtual int fetch_stack3(const DevDispatch& d, int* host3) = 0;  // copy to the host
  virtual int store_stack3(const DevDispatch& d, const int* host3) = 0;
  // 7-wide entries (m,n,k,a,b,c,c_blk) of a dispatch entry in traversal order = what the host builder hands to the scheduler
  virtual int fetch_params7(const DevDispatch& d, int* host7) = 0;
  // new part of the C index [from, from + n)
  virtual int fetch_index(int from, int n, int* rows, int* cols, int* blk_p) = 0;
  // does this stack get its device order on the device?
  static bool device_ordered(const Config& cfg, const StackDescr& d) {
    if (!cfg.stack_sort) return true;  // no ordering at all: traversal order is the device order
    return d.defined_mnk && 2LL * d.max_m * d.max_n * d.max_k > cfg.min_flop_sort;
  }
};

// cuda_stream: cudaStream_t of the owning thread (every pass is enqueued there)
IDeviceBuilder* make_device_builder(void* cuda_stream);
// the same passes in host loops -- CPU tests only
IDeviceBuilder* make_emulated_builder();

}  // namespace dbcsr_b200
