// dbcsr_b200/csrc/host/stack_builder.hpp -- host-side local multiply: recursive index sort, recursive (M,K,N) bisection,
// CSR leaf traversal with C-block creation, (m,n,k)-binned parameter stacks and their dispatch order.
//
// In a real DBCSR build this work is done by the Fortran layer that stays in place
//   (src/mm/dbcsr_mm_common.F:227-309, src/mm/dbcsr_mm_multrec.F:263-658, src/mm/dbcsr_mm_csr.F:178-795,
//    src/mm/dbcsr_mm_accdrv.F:364-423, src/dist/dbcsr_dist_util.F:753-812).
// This C++ implementation exists because the benchmark / test harness needs the same stacks without a Fortran compiler, and as
// the basis of the multi-threaded builder.  It reproduces the reference's traversal so that the C block
// index (order of first touch), the stack contents and the stack dispatch order are identical for one thread.
// All block coordinates and element offsets are 1-based as in DBCSR.
#pragma once
#include <cstdint>
#include <functional>
#include <vector>

namespace dbcsr_b200 {

struct Config {
  int mm_stack_size = 30000;       // DBCSR_MM_STACK_SIZE on accelerator builds (src/core/dbcsr_config.F:76-82)
  int n_stacks = 3;                // DBCSR_N_STACKS: n^3 homogeneous stacks + 1 default
  int multrec_limit = 512;         // DBCSR_MULTREC_LIMIT
  int stack_sort = 1;              // DBCSR_ACCDRV_STACK_SORT
  int min_flop_sort = 4000;        // DBCSR_ACCDRV_MIN_FLOP_SORT
  int binning_nbins = 4096;        // DBCSR_ACCDRV_BINNING_NBINS
  int binning_binsize = 16;        // DBCSR_ACCDRV_BINNING_BINSIZE
};

struct Idx3 {
  int row, col, blk;  // (local row, local col, 1-based element offset)
};

struct StackDescr {
  int m = 0, n = 0, k = 0, max_m = 0, max_n = 0, max_k = 0;
  int defined_mnk = 0;
  // smallest / largest C block id among the entries currently in the stack (maintained by LocalMultiply; 0 = unknown)
  int id_lo = 0, id_hi = 0;
};

// rec_sort_index (src/mm/dbcsr_mm_common.F:227-309): quadtree-like ordering of a block list, in place.
void rec_sort_index(int mi, int mf, int ni, int nf, Idx3* a, int nele, std::vector<Idx3>& tmp);
// same result; the top `depth` levels of the recursion sort their two halves concurrently (2^depth threads at most)
void rec_sort_index_mt(int mi, int mf, int ni, int nf, Idx3* a, int nele, int depth);

// stack_sort / stack_binning (src/mm/dbcsr_mm_accdrv.F:364-423): 7-wide host stack -> 3-wide device stack.
void stack_sort(const int* params7, int* out3, int stack_size);
void stack_binning(const int* params7, int* out3, int stack_size, int nbins, int binsize);
// what dbcsr_mm_accdrv_process does to a stack before upload (:481-491)
// ids_monotone: the caller guarantees that c_first increases strictly with the C block id (column 7) - LocalMultiply's stacks do -
// which allows a one-pass counting sort with the same result as the stable sort by c_first
void accdrv_order_stack(const Config& cfg, const StackDescr& d, const int* params7, int* out3, int stack_size, bool ids_monotone = false);
bool stack_sort_by_block_id(const int* params7, int* out3, int stack_size, int id_lo = 0, int id_hi = 0);

// map_most_common (src/dist/dbcsr_dist_util.F:753-812)
void map_most_common(const std::vector<int>& array, int nmost_common, std::vector<int>& map, std::vector<int>& elements, int& max_val);

class LocalMultiply {
 public:
  // called for every dispatched stack, in dispatch order: (stack number 1-based, descriptor, 7-wide params, size)
  using DispatchFn = std::function<void(int, const StackDescr&, const int*, int)>;

  LocalMultiply(const Config& cfg, const std::vector<int>& m_sizes, const std::vector<int>& n_sizes, const std::vector<int>& k_sizes);

  // Existing C blocks (beta != 0 / retain_sparsity flows): row, col, 1-based offsets; defines datasize
  // (the work matrix starts from the existing blocks: fill_hash_tables, src/mm/dbcsr_mm_csr.F:540-576).
  void preset_c(const int* rows, const int* cols, const int* blk_p, int nblks, int datasize);
  // retain_sparsity of dbcsr_multiply: products whose C block does not exist yet are skipped (src/mm/dbcsr_mm_csr.F:307)
  void set_keep_sparsity(bool keep) { keep_sparsity_ = keep; }
  // Product matrix with symmetry (src/mm/dbcsr_mm_csr.F:280-292): of every off-diagonal pair {(r,c),(c,r)} only the block whose
  // GLOBAL coordinates do not need a transpose under the checkerboard rule (checker_tr, src/dist/dbcsr_dist_operations.F:65-75)
  // is computed.  global_rows / global_cols map the local C rows / cols to global block indices (empty = identity).
  // Ends with the next reset(), like keep_sparsity.
  void set_c_symmetry(bool on, const std::vector<int>& global_rows = {}, const std::vector<int>& global_cols = {}) {
    c_sym_ = on;
    c_grow_ = global_rows;
    c_gcol_ = global_cols;
  }
  static bool checker_tr(int row, int col) { return (((row + col) & 1) != 0) == (col >= row); }

  // One Cannon tick: lists must already be rec-sorted (use sort_panel); [a_first, a_last] = this thread's slice of the
  // left list (1-based, inclusive; the whole list for one thread).  Stacks still partially filled at the end are purged
  // (dbcsr_mm_multrec_multiply -> dbcsr_mm_csr_purge_stacks, src/mm/dbcsr_mm_multrec.F:263-324).
  // a_norms / b_norms (optional, both or neither): per-block norms (sum of squares, single precision, c_calculate_norms /
  // calculate_norms src/mm/dbcsr_mm_common.F:629-670) aligned with a_index / b_index; with set_row_eps they switch on the
  // on-the-fly filter of the CSR leaf (src/mm/dbcsr_mm_csr.F:270-278): product skipped when a_norm*b_norm < row_eps(a_row).
  void multiply(const Idx3* a_index, int a_first, int a_last, const Idx3* b_index, int nb, const DispatchFn& dispatch,
                const float* a_norms = nullptr, const float* b_norms = nullptr);

  static void sort_panel(std::vector<Idx3>& list, int nrows, int ncols);

  // Cannon ticks bring panels with different k-slices: block sizes of the contraction index of the CURRENT panels
  // (DBCSR passes k_sizes per dbcsr_mm_multrec_multiply call, src/mm/dbcsr_mm_multrec.F:263-296; the stack map stays as built).
  void set_k_sizes(const std::vector<int>& k_sizes) { k_sizes_ = k_sizes; }
  int k_size(int kblk_1based) const { return (kblk_1based >= 1 && kblk_1based <= (int)k_sizes_.size()) ? k_sizes_[(size_t)kblk_1based - 1] : 0; }

  // per-C-row threshold row_max_epss (src/mm/dbcsr_mm_cannon.F:1100-1107); empty = no filtering
  void set_row_eps(const std::vector<float>& eps) { row_eps_ = eps; }

  // Forget the product index (new multiply) but keep every allocation (stack buffers, hash table, index capacity).
  void reset();

  // product work matrix (pre-finalize index, order of first touch; src/mm/dbcsr_mm_csr.F:309-323)
  const std::vector<int>& c_row() const { return c_row_; }
  const std::vector<int>& c_col() const { return c_col_; }
  const std::vector<int>& c_blk_p() const { return c_blk_p_; }
  int datasize() const { return datasize_; }
  int64_t flop() const { return flop_; }
  int nstacks() const { return nstacks_; }
  const StackDescr& descr(int istack) const { return descr_[istack]; }  // 1-based

  // ---- hooks of the device-side builder (device_builder.hpp) -------------------------------------------
  // A leaf of sparse_multrec: the block-list ranges [ai,af] x [bi,bf] (1-based, inclusive) handed to the CSR multiply.
  struct Leaf {
    int ai, af, bi, bf;
  };
  // The recursion of multiply() without its leaves: the same cuts in the same order, the leaves appended to `out`.
  void plan(const Idx3* a_index, int a_first, int a_last, const Idx3* b_index, int nb, std::vector<Leaf>& out);
  // Blocks created outside csr_multiply_low (by the device builder, in first-touch order) become part of the work index; the
  // per-row lookup tables are NOT updated: a multiply uses either the host or the device builder from reset() to reset().
  void append_index(const int* rows, const int* cols, const int* blk_p, int nblks, int datasize, int64_t flop_add);
  const std::vector<int>& m_sizes() const { return m_sizes_; }
  const std::vector<int>& n_sizes() const { return n_sizes_; }
  const std::vector<int>& k_sizes() const { return k_sizes_; }
  const std::vector<int>& m_map() const { return m_map_; }
  const std::vector<int>& n_map() const { return n_map_; }
  const std::vector<int>& k_map() const { return k_map_; }
  const std::vector<int>& stack_map() const { return stack_map_; }
  const Config& config() const { return cfg_; }
  bool keep_sparsity() const { return keep_sparsity_; }
  bool c_symmetry() const { return c_sym_; }
  const std::vector<int>& c_global_rows() const { return c_grow_; }
  const std::vector<int>& c_global_cols() const { return c_gcol_; }

 private:
  void init_stack_map();
  void sparse_multrec(int mi, int mf, int ni, int nf, int ki, int kf, int ai, int af, const Idx3* a, int bi, int bf, const Idx3* b);
  void csr_multiply_low(int mi, int mf, int ki, int kf, int ai, int af, int bi, int bf, const Idx3* a, const Idx3* b);
  void flush_stacks(bool purge);
  int c_lookup_or_insert(int row, int col, int nze, bool& created, bool insert = true);  // 0: absent and !insert

  Config cfg_;
  std::vector<int> m_sizes_, n_sizes_, k_sizes_;
  std::vector<int> m_map_, n_map_, k_map_;
  int max_m_ = 0, max_n_ = 0, max_k_ = 0;
  int nstacks_ = 0;
  std::vector<int> stack_map_;  // [(m_map-1)*(n+1)*(n+1) + (k_map-1)*(n+1) + (n_map-1)] -> stack number
  std::vector<StackDescr> descr_;
  std::vector<std::vector<int>> stacks_;  // 7 ints per entry
  std::vector<int> fill_;
  const DispatchFn* dispatch_ = nullptr;
  // C index + one small open-addressing table per C row (col -> c_blk_id), like the reference's c_hashes(a_row)
  // (src/mm/dbcsr_mm_csr.F:294, src/utils/dbcsr_hash_table.f90): row-local tables stay cache resident during a CSR leaf
  struct RowTable {
    std::vector<int> cols, ids;  // ids == 0: empty slot
    int mask = 0, count = 0;
    std::vector<int> dense;      // direct table col -> id (0 = absent) when the matrix has few enough block columns
  };
  static constexpr int kDenseRowLimit = 4096;  // block columns up to which a touched C row gets a direct table (16 KB per row, 64 MB at most)
  std::vector<int> c_row_, c_col_, c_blk_p_;
  std::vector<RowTable> rows_;
  int datasize_ = 0;
  int64_t flop_ = 0;
  bool keep_sparsity_ = false;
  bool c_sym_ = false;
  std::vector<int> c_grow_, c_gcol_;
  // scratch for the CSR leaves
  std::vector<int> a_row_p_, b_row_p_, a_info_, b_info_, counts_;
  // on-the-fly filter (optional)
  std::vector<float> row_eps_, a_csr_norms_, b_csr_norms_;
  const float* a_norms_ = nullptr;
  const float* b_norms_ = nullptr;
  std::vector<Leaf>* plan_out_ = nullptr;
};

}  // namespace dbcsr_b200
