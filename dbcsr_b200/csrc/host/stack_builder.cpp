// dbcsr_b200/csrc/host/stack_builder.cpp -- see stack_builder.hpp.  New C++ code following the reference's traversal rules.
#include "stack_builder.hpp"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <thread>

namespace dbcsr_b200 {

// ------------------------------------------------------------------------------------------------ rec_sort_index
// Reference: rec_split fills the low part front-to-back and the high part back-to-front (so the high part comes out
// reversed), then recurses into each part while it holds more than one element (src/mm/dbcsr_mm_common.F:248-303).
// Always halves the longer extent; on a tie the columns.
void rec_sort_index(int mi, int mf, int ni, int nf, Idx3* a, int nele, std::vector<Idx3>& tmp) {
  if (nele <= 0) return;
  const int M = mf - mi + 1, N = nf - ni + 1;
  if ((int)tmp.size() < nele) tmp.resize(nele);
  const bool by_row = M > N;
  const int half = by_row ? M / 2 : N / 2;
  const int half_m = (by_row ? mi : ni) + half - 1;
  int p_low = 0, p_high = nele - 1;
  for (int el = 0; el < nele; ++el) {
    const int key = by_row ? a[el].row : a[el].col;
    if (key <= half_m)
      tmp[p_low++] = a[el];
    else
      tmp[p_high--] = a[el];
  }
  std::memcpy(a, tmp.data(), sizeof(Idx3) * (size_t)nele);
  const int nlow = p_low;
  if (by_row) {
    if (nlow > 1) rec_sort_index(mi, mi + half - 1, ni, nf, a, nlow, tmp);
    if (nele - nlow > 1) rec_sort_index(mi + half, mf, ni, nf, a + nlow, nele - nlow, tmp);
  }
  else {
    if (nlow > 1) rec_sort_index(mi, mf, ni, ni + half - 1, a, nlow, tmp);
    if (nele - nlow > 1) rec_sort_index(mi, mf, ni + half, nf, a + nlow, nele - nlow, tmp);
  }
}

// The two halves produced by one split are independent: the top `depth` levels of the recursion run them concurrently (the
// right panel of a multiply is sorted as one list of ~1e5 blocks on the critical path of every multiply; the order is the same).
void rec_sort_index_mt(int mi, int mf, int ni, int nf, Idx3* a, int nele, int depth) {
  if (nele <= 1) return;
  if (depth <= 0 || nele < 8192) {
    std::vector<Idx3> tmp;
    rec_sort_index(mi, mf, ni, nf, a, nele, tmp);
    return;
  }
  const int M = mf - mi + 1, N = nf - ni + 1;
  const bool by_row = M > N;
  const int half = by_row ? M / 2 : N / 2;
  const int half_m = (by_row ? mi : ni) + half - 1;
  int nlow = 0;
  {
    std::vector<Idx3> tmp((size_t)nele);
    int p_low = 0, p_high = nele - 1;
    for (int el = 0; el < nele; ++el) {
      const int key = by_row ? a[el].row : a[el].col;
      if (key <= half_m)
        tmp[(size_t)p_low++] = a[el];
      else
        tmp[(size_t)p_high--] = a[el];
    }
    std::memcpy(a, tmp.data(), sizeof(Idx3) * (size_t)nele);
    nlow = p_low;
  }
  const int lo_mf = by_row ? mi + half - 1 : mf, lo_nf = by_row ? nf : ni + half - 1;
  const int hi_mi = by_row ? mi + half : mi, hi_ni = by_row ? ni : ni + half;
  std::thread low([=]() { rec_sort_index_mt(mi, lo_mf, ni, lo_nf, a, nlow, depth - 1); });
  rec_sort_index_mt(hi_mi, mf, hi_ni, nf, a + nlow, nele - nlow, depth - 1);
  low.join();
}

void LocalMultiply::sort_panel(std::vector<Idx3>& list, int nrows, int ncols) {
  std::vector<Idx3> tmp(list.size());
  if (!list.empty()) rec_sort_index(1, nrows, 1, ncols, list.data(), (int)list.size(), tmp);
}

// ------------------------------------------------------------------------------------------------ stack ordering
void stack_sort(const int* params7, int* out3, int stack_size) {
  // DBCSR's sort() is a stable merge sort (src/utils/dbcsr_array_sort.F) by c_first (params(6,:)).  Same result with a stable
  // LSD radix sort (3 passes of 11 bits over the non-negative 32-bit keys): the host-side sort is on the critical path of every
  // stack (src/mm/dbcsr_mm_accdrv.F:476-478 flags it), a comparison sort of 30000 keys costs more than building the stack.
  if (stack_size <= 0) return;
  static thread_local std::vector<uint32_t> idx_a, idx_b;
  idx_a.resize((size_t)stack_size);
  idx_b.resize((size_t)stack_size);
  uint32_t kmax = 0;
  for (int i = 0; i < stack_size; ++i) {
    idx_a[(size_t)i] = (uint32_t)i;
    kmax |= (uint32_t)params7[7 * (size_t)i + 5];
  }
  uint32_t* src = idx_a.data();
  uint32_t* dst = idx_b.data();
  for (int shift = 0; shift < 32 && (kmax >> shift) != 0; shift += 11) {
    uint32_t count[2049] = {0};
    for (int i = 0; i < stack_size; ++i) count[(((uint32_t)params7[7 * (size_t)src[i] + 5] >> shift) & 2047u) + 1]++;
    for (int d = 0; d < 2048; ++d) count[d + 1] += count[d];
    for (int i = 0; i < stack_size; ++i) {
      const uint32_t d = ((uint32_t)params7[7 * (size_t)src[i] + 5] >> shift) & 2047u;
      dst[count[d]++] = src[i];
    }
    std::swap(src, dst);
  }
  for (int i = 0; i < stack_size; ++i) {
    const int* p = params7 + 7 * (size_t)src[i];
    out3[3 * (size_t)i] = p[3];
    out3[3 * (size_t)i + 1] = p[4];
    out3[3 * (size_t)i + 2] = p[5];
  }
}

// Same order as stack_sort when c_first is strictly increasing in the C block id params(7,:) over the entries of the stack --
// true for every stack LocalMultiply builds (a new block gets offset datasize + 1, existing blocks were laid out in index order):
// ONE counting-sort pass over the (narrow) id range of the stack instead of three radix passes over random 28-byte records.
bool stack_sort_by_block_id(const int* params7, int* out3, int stack_size, int id_lo, int id_hi) {
  if (stack_size <= 0) return true;
  int lo = id_lo, hi = id_hi;
  if (lo <= 0 || hi < lo) {  // range not tracked by the builder: one more pass
    lo = hi = params7[6];
    for (int i = 1; i < stack_size; ++i) {
      const int id = params7[7 * (size_t)i + 6];
      lo = std::min(lo, id);
      hi = std::max(hi, id);
    }
  }
  const long long range = (long long)hi - lo + 1;
  if (range > 4LL * stack_size + 4096) {
    // ids scattered (the stack revisits old C tiles): stable LSD radix sort of compact (key, position) pairs, 11 bits per pass --
    // two passes cover 4M blocks -- then one gather; the 28-byte records are read once sequentially and once by position
    static thread_local std::vector<uint64_t> pa, pb;
    pa.resize((size_t)stack_size);
    pb.resize((size_t)stack_size);
    for (int i = 0; i < stack_size; ++i) pa[(size_t)i] = ((uint64_t)(uint32_t)(params7[7 * (size_t)i + 6] - lo) << 32) | (uint32_t)i;
    uint64_t* src = pa.data();
    uint64_t* dst = pb.data();
    for (int shift = 0; shift < 32 && ((uint64_t)(range - 1) >> shift) != 0; shift += 11) {
      uint32_t cnt[2049] = {0};
      for (int i = 0; i < stack_size; ++i) cnt[((src[i] >> (32 + shift)) & 2047u) + 1]++;
      for (int d = 0; d < 2048; ++d) cnt[d + 1] += cnt[d];
      for (int i = 0; i < stack_size; ++i) dst[cnt[(src[i] >> (32 + shift)) & 2047u]++] = src[i];
      std::swap(src, dst);
    }
    for (int i = 0; i < stack_size; ++i) {
      const int* p = params7 + 7 * (size_t)(uint32_t)src[i];
      out3[3 * (size_t)i] = p[3];
      out3[3 * (size_t)i + 1] = p[4];
      out3[3 * (size_t)i + 2] = p[5];
    }
    return true;
  }
  static thread_local std::vector<int> count;
  count.assign((size_t)range + 1, 0);
  for (int i = 0; i < stack_size; ++i) count[(size_t)(params7[7 * (size_t)i + 6] - lo) + 1]++;
  for (long long d = 0; d < range; ++d) count[(size_t)d + 1] += count[(size_t)d];
  for (int i = 0; i < stack_size; ++i) {
    const int* p = params7 + 7 * (size_t)i;
    const int pos = count[(size_t)(p[6] - lo)]++;
    out3[3 * (size_t)pos] = p[3];
    out3[3 * (size_t)pos + 1] = p[4];
    out3[3 * (size_t)pos + 2] = p[5];
  }
  return true;
}

void stack_binning(const int* params7, int* out3, int stack_size, int nbins, int binsize) {
  std::vector<int> bin_arr((size_t)3 * binsize * nbins);
  std::vector<int> bin_top((size_t)nbins, 0);
  size_t top = 0;
  for (int i = 0; i < stack_size; ++i) {
    const int* val = params7 + 7 * (size_t)i + 3;
    // src/mm/dbcsr_mm_accdrv.F:405-406: INT(val(3)*(val(3)+3), KIND=int_8) multiplies in 32-bit INTEGER before widening, so
    // the product wraps (two's complement) for c_first > 46339; MODULO floors => bin id in [0, nbins)
    const int32_t prod = (int32_t)((uint32_t)val[2] * (uint32_t)(val[2] + 3));
    const int bin_id = (int)((((int64_t)prod % nbins) + nbins) % nbins);
    if (bin_top[bin_id] >= binsize) {
      std::memcpy(out3 + 3 * top, &bin_arr[(size_t)3 * binsize * bin_id], sizeof(int) * 3 * (size_t)bin_top[bin_id]);
      top += bin_top[bin_id];
      bin_top[bin_id] = 0;
    }
    int* dst = &bin_arr[(size_t)3 * binsize * bin_id + 3 * (size_t)bin_top[bin_id]];
    dst[0] = val[0];
    dst[1] = val[1];
    dst[2] = val[2];
    bin_top[bin_id]++;
  }
  for (int b = 0; b < nbins; ++b) {
    if (bin_top[b] > 0) {
      std::memcpy(out3 + 3 * top, &bin_arr[(size_t)3 * binsize * b], sizeof(int) * 3 * (size_t)bin_top[b]);
      top += bin_top[b];
    }
  }
}

void accdrv_order_stack(const Config& cfg, const StackDescr& d, const int* params7, int* out3, int stack_size, bool ids_monotone) {
  const int64_t flop_per_entry = 2LL * d.max_m * d.max_n * d.max_k;
  if (cfg.stack_sort) {
    if (flop_per_entry > cfg.min_flop_sort) {
      if (!(ids_monotone && stack_sort_by_block_id(params7, out3, stack_size, d.id_lo, d.id_hi))) stack_sort(params7, out3, stack_size);
    }
    else
      stack_binning(params7, out3, stack_size, cfg.binning_nbins, cfg.binning_binsize);
  }
  else {
    for (int i = 0; i < stack_size; ++i) {
      out3[3 * (size_t)i] = params7[7 * (size_t)i + 3];
      out3[3 * (size_t)i + 1] = params7[7 * (size_t)i + 4];
      out3[3 * (size_t)i + 2] = params7[7 * (size_t)i + 5];
    }
  }
}

// ------------------------------------------------------------------------------------------------ map_most_common
void map_most_common(const std::vector<int>& array, int nmost_common, std::vector<int>& map, std::vector<int>& elements, int& max_val) {
  int max_val_l = 0;
  max_val = 0;
  if (!array.empty()) {
    max_val = *std::max_element(array.begin(), array.end());
    max_val_l = max_val;
  }
  std::vector<int> size_counts((size_t)max_val_l + 1, 0), perm((size_t)max_val_l + 1, 0);
  for (int v : array)
    if (v <= max_val_l) size_counts[v] -= 1;  // negative counts: ascending stable sort = most frequent first
  if (!array.empty()) {
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](int x, int y) { return size_counts[x] < size_counts[y]; });
  }
  const int nmc = std::min(nmost_common, max_val_l);
  map.assign((size_t)max_val_l + 1, nmost_common + 1);
  for (int i = 1; i <= nmc; ++i) map[perm[i - 1]] = i;
  elements.assign((size_t)nmost_common, 0);
  for (int i = 0; i < nmc; ++i) elements[i] = perm[i];
}

// ------------------------------------------------------------------------------------------------ LocalMultiply
LocalMultiply::LocalMultiply(const Config& cfg, const std::vector<int>& m_sizes, const std::vector<int>& n_sizes,
                             const std::vector<int>& k_sizes)
    : cfg_(cfg), m_sizes_(m_sizes), n_sizes_(n_sizes), k_sizes_(k_sizes) {
  init_stack_map();
  stacks_.resize((size_t)nstacks_ + 1);
  fill_.assign((size_t)nstacks_ + 1, 0);
  // stack buffers (7 x mm_stack_size ints each) are allocated at first use: most of the n^3+1 stacks never see an entry
  rows_.resize(m_sizes_.size() + 1);
}

// src/mm/dbcsr_mm_csr.F:404-525
void LocalMultiply::init_stack_map() {
  const int ns = cfg_.n_stacks;
  nstacks_ = ns * ns * ns + 1;
  std::vector<int> mc_m, mc_n, mc_k;
  map_most_common(m_sizes_, ns, m_map_, mc_m, max_m_);
  map_most_common(n_sizes_, ns, n_map_, mc_n, max_n_);
  map_most_common(k_sizes_, ns, k_map_, mc_k, max_k_);
  const int w = ns + 1;
  stack_map_.assign((size_t)w * w * w, nstacks_);
  std::vector<StackDescr> descr((size_t)nstacks_ + 1);
  for (int m_map = 1; m_map <= ns + 1; ++m_map)
    for (int k_map = 1; k_map <= ns + 1; ++k_map)
      for (int n_map = 1; n_map <= ns + 1; ++n_map) {
        const size_t slot = ((size_t)(m_map - 1) * w + (k_map - 1)) * w + (n_map - 1);
        if (m_map <= ns && k_map <= ns && n_map <= ns) {
          int ps_g = (m_map - 1) * ns * ns + (k_map - 1) * ns + n_map;
          ps_g = nstacks_ - ps_g;
          stack_map_[slot] = ps_g;
          StackDescr& d = descr[ps_g];
          d.m = d.max_m = mc_m[m_map - 1];
          d.n = d.max_n = mc_n[n_map - 1];
          d.k = d.max_k = mc_k[k_map - 1];
          d.defined_mnk = 1;
        }
        else {
          stack_map_[slot] = nstacks_;
          StackDescr& d = descr[nstacks_];
          d.m = d.n = d.k = 0;
          d.max_m = max_m_;
          d.max_n = max_n_;
          d.max_k = max_k_;
          d.defined_mnk = 0;
        }
      }
  // order the homogeneous stacks by decreasing flops (stable), default stack stays last
  std::vector<int> order((size_t)nstacks_ - 1);
  std::iota(order.begin(), order.end(), 1);
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
    return -2LL * descr[x].m * descr[x].n * descr[x].k < -2LL * descr[y].m * descr[y].n * descr[y].k;
  });
  descr_.assign((size_t)nstacks_ + 1, StackDescr());
  std::vector<int> newpos((size_t)nstacks_ + 1, 0);
  for (int i = 1; i < nstacks_; ++i) {
    descr_[i] = descr[order[i - 1]];
    newpos[order[i - 1]] = i;
  }
  descr_[nstacks_] = descr[nstacks_];
  newpos[nstacks_] = nstacks_;
  for (auto& s : stack_map_) s = newpos[s];
}

void LocalMultiply::reset() {
  c_row_.clear();
  c_col_.clear();
  c_blk_p_.clear();
  for (auto& t : rows_) {
    if (t.count > 0) {
      std::fill(t.ids.begin(), t.ids.end(), 0);
      std::fill(t.dense.begin(), t.dense.end(), 0);
    }
    t.count = 0;
  }
  datasize_ = 0;
  flop_ = 0;
  keep_sparsity_ = false;
  c_sym_ = false;
  c_grow_.clear();
  c_gcol_.clear();
  std::fill(fill_.begin(), fill_.end(), 0);
}

void LocalMultiply::preset_c(const int* rows, const int* cols, const int* blk_p, int nblks, int datasize) {
  for (int i = 0; i < nblks; ++i) {
    bool created = false;
    const int id = c_lookup_or_insert(rows[i], cols[i], 0, created);
    c_blk_p_[id - 1] = blk_p[i];
  }
  datasize_ = datasize;
}

// hash_table_get / hash_table_add of the reference (src/utils/dbcsr_hash_table.f90) only affect speed, not results;
// new block => offset = datasize + 1, appended to the work index (src/mm/dbcsr_mm_csr.F:309-323).
int LocalMultiply::c_lookup_or_insert(int row, int col, int nze, bool& created, bool insert) {
  RowTable& t = rows_[(size_t)row];
  if ((int)n_sizes_.size() <= kDenseRowLimit) {
    // few block columns: one direct table per touched row (row-local and cache resident during a CSR leaf, like the hash table
    // it replaces; the reference's hash only affects speed, src/utils/dbcsr_hash_table.f90)
    if (t.dense.empty()) t.dense.assign(n_sizes_.size() + 1, 0);
    int& slot = t.dense[(size_t)col];
    if (slot != 0) {
      created = false;
      return slot;
    }
    if (!insert) {
      created = false;
      return 0;
    }
    created = true;
    c_row_.push_back(row);
    c_col_.push_back(col);
    c_blk_p_.push_back(datasize_ + 1);
    datasize_ += nze;
    slot = (int)c_blk_p_.size();
    ++t.count;
    return slot;
  }
  if (t.mask == 0) {
    t.cols.assign(16, 0);
    t.ids.assign(16, 0);
    t.mask = 15;
  }
  unsigned p = ((unsigned)col * 2654435761u >> 7) & (unsigned)t.mask;
  while (t.ids[p] != 0) {
    if (t.cols[p] == col) {
      created = false;
      return t.ids[p];
    }
    p = (p + 1) & (unsigned)t.mask;
  }
  if (!insert) {
    created = false;
    return 0;
  }
  created = true;
  c_row_.push_back(row);
  c_col_.push_back(col);
  c_blk_p_.push_back(datasize_ + 1);
  datasize_ += nze;
  const int id = (int)c_blk_p_.size();
  t.cols[p] = col;
  t.ids[p] = id;
  if (++t.count * 2 > t.mask) {  // grow x4 and re-insert
    std::vector<int> oc, oi;
    oc.swap(t.cols);
    oi.swap(t.ids);
    const int ncap = (t.mask + 1) * 4;
    t.cols.assign((size_t)ncap, 0);
    t.ids.assign((size_t)ncap, 0);
    t.mask = ncap - 1;
    for (size_t i = 0; i < oi.size(); ++i) {
      if (oi[i] != 0) {
        unsigned q = ((unsigned)oc[i] * 2654435761u >> 7) & (unsigned)t.mask;
        while (t.ids[q] != 0) q = (q + 1) & (unsigned)t.mask;
        t.cols[q] = oc[i];
        t.ids[q] = oi[i];
      }
    }
  }
  return id;
}

// src/mm/dbcsr_mm_csr.F:704-739 (+ sched/accdrv hand-off through the dispatch call-back)
void LocalMultiply::flush_stacks(bool purge) {
  const int min_fill = purge ? 0 : cfg_.mm_stack_size * 3 / 4;
  for (int i = 1; i <= nstacks_; ++i) {
    if (fill_[i] > min_fill) {
      (*dispatch_)(i, descr_[i], stacks_[i].data(), fill_[i]);
      fill_[i] = 0;
    }
  }
}

// src/mm/dbcsr_mm_csr.F:178-359 with build_csr_index :741-795 (optional norm filter, optional symmetric-product skipping)
void LocalMultiply::csr_multiply_low(int mi, int mf, int ki, int kf, int ai, int af, int bi, int bf, const Idx3* a, const Idx3* b) {
  const int nrow = mf - mi + 1, nk = kf - ki + 1, na = af - ai + 1, nb = bf - bi + 1;
  const bool use_eps = a_norms_ != nullptr && b_norms_ != nullptr && !row_eps_.empty();
  auto build = [&](int lo, int n_rows, int first, int count, const Idx3* lst, std::vector<int>& row_p, std::vector<int>& info,
                   const float* list_norms, std::vector<float>& csr_norms) {
    row_p.assign((size_t)n_rows + 1, 0);
    counts_.assign((size_t)n_rows, 0);
    for (int i = first; i < first + count; ++i) counts_[lst[i - 1].row - lo]++;
    for (int r = 1; r <= n_rows; ++r) row_p[r] = row_p[r - 1] + counts_[r - 1];
    info.resize((size_t)2 * count);
    if (use_eps) csr_norms.resize((size_t)count);
    std::fill(counts_.begin(), counts_.end(), 0);
    for (int i = first; i < first + count; ++i) {
      const int r = lst[i - 1].row - lo;
      const int pos = row_p[r] + counts_[r]++;
      info[2 * (size_t)pos] = lst[i - 1].col;
      info[2 * (size_t)pos + 1] = lst[i - 1].blk;
      if (use_eps) csr_norms[(size_t)pos] = list_norms[i - 1];
    }
  };
  build(mi, nrow, ai, na, a, a_row_p_, a_info_, a_norms_, a_csr_norms_);
  build(ki, nk, bi, nb, b, b_row_p_, b_info_, b_norms_, b_csr_norms_);
  const int w = cfg_.n_stacks + 1;
  for (int a_row_l = mi; a_row_l <= mf; ++a_row_l) {
    const int m_size = m_sizes_[a_row_l - 1];
    const int mapped_row = m_size < (int)m_map_.size() ? m_map_[m_size] : cfg_.n_stacks + 1;
    const float a_row_eps = use_eps ? row_eps_[(size_t)a_row_l - 1] : 0.0f;
    for (int a_blk = a_row_p_[a_row_l - mi]; a_blk < a_row_p_[a_row_l - mi + 1]; ++a_blk) {
      const int a_col_l = a_info_[2 * (size_t)a_blk];
      const int a_first = a_info_[2 * (size_t)a_blk + 1];
      const int k_size = k_sizes_[a_col_l - 1];
      const int mapped_k = k_size < (int)k_map_.size() ? k_map_[k_size] : cfg_.n_stacks + 1;
      const float a_norm = use_eps ? a_csr_norms_[(size_t)a_blk] : 0.0f;
      for (int b_blk = b_row_p_[a_col_l - ki]; b_blk < b_row_p_[a_col_l - ki + 1]; ++b_blk) {
        if (use_eps) {  // single-precision product and compare, like the reference
          const float prod = a_norm * b_csr_norms_[(size_t)b_blk];
          if (prod < a_row_eps) continue;
        }
        const int b_col_l = b_info_[2 * (size_t)b_blk];
        if (c_sym_) {  // don't calculate symmetric blocks (src/mm/dbcsr_mm_csr.F:280-292)
          const int cr = c_grow_.empty() ? a_row_l : c_grow_[(size_t)a_row_l - 1];
          const int cc = c_gcol_.empty() ? b_col_l : c_gcol_[(size_t)b_col_l - 1];
          if (cr != cc && checker_tr(cr, cc)) continue;
        }
        const int b_first = b_info_[2 * (size_t)b_blk + 1];
        const int n_size = n_sizes_[b_col_l - 1];
        const int c_nze = m_size * n_size;
        bool created = false;
        const int c_blk_id = c_lookup_or_insert(a_row_l, b_col_l, c_nze, created, !keep_sparsity_);
        if (c_blk_id == 0) continue;  // keep_sparsity: no new blocks
        const int offset = c_blk_p_[c_blk_id - 1];
        // zero-sized blocks (DBCSR allows block size 0): the C block exists from here on (src/mm/dbcsr_mm_csr.F:325-333 "we
        // still need to get to here to get new blocks"), but there is nothing to multiply: no stack entry (kernels need m,n,k >= 1)
        if (c_nze == 0 || k_size == 0) continue;
        const int mapped_col = n_size < (int)n_map_.size() ? n_map_[n_size] : cfg_.n_stacks + 1;
        const int ws = stack_map_[((size_t)(mapped_row - 1) * w + (mapped_k - 1)) * w + (mapped_col - 1)];
        if (stacks_[ws].empty()) stacks_[ws].resize((size_t)7 * cfg_.mm_stack_size);
        int* p = stacks_[ws].data() + 7 * (size_t)fill_[ws];
        p[0] = m_size;
        p[1] = n_size;
        p[2] = k_size;
        p[3] = a_first;
        p[4] = b_first;
        p[5] = offset;
        p[6] = c_blk_id;
        StackDescr& sd = descr_[ws];
        if (fill_[ws] == 0) {
          sd.id_lo = sd.id_hi = c_blk_id;
        }
        else {
          sd.id_lo = std::min(sd.id_lo, c_blk_id);
          sd.id_hi = std::max(sd.id_hi, c_blk_id);
        }
        fill_[ws]++;
        flop_ += 2LL * c_nze * k_size;
        if (fill_[ws] >= cfg_.mm_stack_size) flush_stacks(false);
      }
    }
  }
}

namespace {
// find_cut_row / find_cut_col, src/mm/dbcsr_mm_multrec.F:579-658 (1-based ai..af)
template <bool BY_ROW>
int find_cut(const Idx3* a, int ai, int af, int val) {
  auto key = [&](int i) { return BY_ROW ? a[i - 1].row : a[i - 1].col; };
  int ilow = ai;
  if (key(ilow) > val) return ilow;
  int ihigh = af;
  if (key(ihigh) <= val) return ihigh + 1;
  for (;;) {
    if (ihigh - ilow == 1) break;
    const int i = (ilow + ihigh) / 2;
    if (key(i) > val)
      ihigh = i;
    else
      ilow = i;
  }
  return ihigh;
}
}  // namespace

// src/mm/dbcsr_mm_multrec.F:487-576: cut the longest of M,K,N (ties: N over K over M) until both lists are short
void LocalMultiply::sparse_multrec(int mi, int mf, int ni, int nf, int ki, int kf, int ai, int af, const Idx3* a, int bi, int bf,
                                   const Idx3* b) {
  if (af < ai || bf < bi || mf < mi || nf < ni || kf < ki) return;
  if (af - ai + 1 <= cfg_.multrec_limit && bf - bi + 1 <= cfg_.multrec_limit) {
    if (plan_out_ != nullptr) {
      plan_out_->push_back(Leaf{ai, af, bi, bf});
      return;
    }
    csr_multiply_low(mi, mf, ki, kf, ai, af, bi, bf, a, b);
    return;
  }
  const int M = mf - mi + 1, N = nf - ni + 1, K = kf - ki + 1;
  int cut = 0;
  if (M >= std::max(N, K)) cut = 1;
  if (K >= std::max(N, M)) cut = 2;
  if (N >= std::max(M, K)) cut = 3;
  if (cut == 1) {
    const int s1 = M / 2;
    const int acut = find_cut<true>(a, ai, af, mi + s1 - 1);
    sparse_multrec(mi, mi + s1 - 1, ni, nf, ki, kf, ai, acut - 1, a, bi, bf, b);
    sparse_multrec(mi + s1, mf, ni, nf, ki, kf, acut, af, a, bi, bf, b);
  }
  else if (cut == 2) {
    const int s1 = K / 2;
    const int acut = find_cut<false>(a, ai, af, ki + s1 - 1);
    const int bcut = find_cut<true>(b, bi, bf, ki + s1 - 1);
    sparse_multrec(mi, mf, ni, nf, ki, ki + s1 - 1, ai, acut - 1, a, bi, bcut - 1, b);
    sparse_multrec(mi, mf, ni, nf, ki + s1, kf, acut, af, a, bcut, bf, b);
  }
  else {
    const int s1 = N / 2;
    const int bcut = find_cut<false>(b, bi, bf, ni + s1 - 1);
    sparse_multrec(mi, mf, ni, ni + s1 - 1, ki, kf, ai, af, a, bi, bcut - 1, b);
    sparse_multrec(mi, mf, ni + s1, nf, ki, kf, ai, af, a, bcut, bf, b);
  }
}

void LocalMultiply::multiply(const Idx3* a_index, int a_first, int a_last, const Idx3* b_index, int nb, const DispatchFn& dispatch,
                             const float* a_norms, const float* b_norms) {
  dispatch_ = &dispatch;
  a_norms_ = a_norms;
  b_norms_ = b_norms;
  sparse_multrec(1, (int)m_sizes_.size(), 1, (int)n_sizes_.size(), 1, (int)k_sizes_.size(), a_first, a_last, a_index, 1, nb, b_index);
  flush_stacks(true);
  dispatch_ = nullptr;
  a_norms_ = b_norms_ = nullptr;
}

void LocalMultiply::plan(const Idx3* a_index, int a_first, int a_last, const Idx3* b_index, int nb, std::vector<Leaf>& out) {
  plan_out_ = &out;
  sparse_multrec(1, (int)m_sizes_.size(), 1, (int)n_sizes_.size(), 1, (int)k_sizes_.size(), a_first, a_last, a_index, 1, nb, b_index);
  plan_out_ = nullptr;
}

void LocalMultiply::append_index(const int* rows, const int* cols, const int* blk_p, int nblks, int datasize, int64_t flop_add) {
  c_row_.insert(c_row_.end(), rows, rows + nblks);
  c_col_.insert(c_col_.end(), cols, cols + nblks);
  c_blk_p_.insert(c_blk_p_.end(), blk_p, blk_p + nblks);
  datasize_ = datasize;
  flop_ += flop_add;
}

}  // namespace dbcsr_b200
