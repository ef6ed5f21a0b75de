// dbcsr_b200/csrc/host/device_builder.cu -- see device_builder.hpp.  Element-wise passes + scans + stable radix sorts that
// reproduce csr_multiply_low / flush_stacks / stack_sort (src/mm/dbcsr_mm_csr.F:178-359, :704-739, src/mm/dbcsr_mm_accdrv.F:364-384)
// for all products of a Cannon tick at once.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "device_builder.hpp"

namespace dbcsr_b200 {
namespace {

#define DB_HD __host__ __device__ __forceinline__

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned char u8;

constexpr u32 kEmpty = 0xFFFFFFFFu;  // table slot without a block
constexpr u32 kNew = 0x80000000u;    // table value = kNew | (traversal position of the earliest product touching the block)
constexpr u32 kNoSlot = 0xFFFFFFFFu; // retain_sparsity: the product's C block does not exist
constexpr int kDropBin = 255;        // products without a stack entry (zero-sized blocks)
constexpr int kMaxBins = 254;
constexpr long long kDenseLimitDefault = 1ll << 25;  // (row, col) slots up to which the C table is a direct array
long long dense_limit() {  // DBCSR_B200_DEVBUILD_DENSE_LIMIT: tests force the open-addressing table on small grids
  const char* s = std::getenv("DBCSR_B200_DEVBUILD_DENSE_LIMIT");
  return s != nullptr ? std::atoll(s) : kDenseLimitDefault;
}

DB_HD u32 atomic_min_u32(u32* p, u32 v) {
#ifdef __CUDA_ARCH__
  return atomicMin(p, v);
#else
  const u32 o = *p;
  if (v < o) *p = v;
  return o;
#endif
}
DB_HD u64 atomic_cas_u64(u64* p, u64 cmp, u64 v) {
#ifdef __CUDA_ARCH__
  return atomicCAS(p, cmp, v);
#else
  const u64 o = *p;
  if (o == cmp) *p = v;
  return o;
#endif
}
template <class T, class V>
DB_HD int lower_bound_n(const T* a, int n, V v) {  // first index with a[i] >= v
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}
template <class T, class V>
DB_HD int upper_bound_n(const T* a, int n, V v) {  // first index with a[i] > v
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= v)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

// ---------------------------------------------------------------------------------------------------------- passes
// build_csr_index of every distinct leaf range (src/mm/dbcsr_mm_csr.F:741-795): stable order by row = rank counting
struct RankRange {
  const int* list;       // (row, col, blk) triples
  const int* rng_start;  // first list position of the range (0-based)
  const int* rng_soff;   // prefix of the range lengths, nranges + 1
  int nranges;
  int* sorted_idx;  // out: list position
  int* sorted_key;  // out: row
  DB_HD void operator()(long long i) const {
    const int r = upper_bound_n(rng_soff, nranges + 1, (int)i) - 1;
    const int j = (int)i - rng_soff[r], len = rng_soff[r + 1] - rng_soff[r];
    const int* base = list + 3 * (size_t)rng_start[r];
    const int key = base[3 * (size_t)j];
    int rank = 0;
    for (int t = 0; t < len; ++t) {
      const int kt = base[3 * (size_t)t];
      rank += (kt < key || (kt == key && t < j)) ? 1 : 0;
    }
    sorted_idx[rng_soff[r] + rank] = rng_start[r] + j;
    sorted_key[rng_soff[r] + rank] = key;
  }
};

// the tests csr_multiply_low applies to a pair before it touches the C index (src/mm/dbcsr_mm_csr.F:270-292)
struct PairTest {
  const float *a_norms = nullptr, *b_norms = nullptr, *row_eps = nullptr;  // by sorted-list position / by block row
  int c_sym = 0;
  const int *grow = nullptr, *gcol = nullptr;
  DB_HD bool active() const { return a_norms != nullptr || c_sym != 0; }
  DB_HD bool keep(int a_pos, int b_pos, int row, int col) const {
    if (a_norms != nullptr) {  // single-precision product and compare, like the reference
      const float prod = a_norms[a_pos] * b_norms[b_pos];
      if (prod < row_eps[row - 1]) return false;
    }
    if (c_sym) {
      const int cr = grow != nullptr ? grow[row - 1] : row, cc = gcol != nullptr ? gcol[col - 1] : col;
      if (cr != cc && ((((cr + cc) & 1) != 0) == (cc >= cr))) return false;  // checker_tr
    }
    return true;
  }
};

// item = (leaf, A block in the leaf's CSR order): the B blocks of row a_col inside the leaf's right range
struct CountItems {
  const int* leaf_ioff;  // prefix of the leaves' A-range lengths, nleaves + 1
  int nleaves;
  const int *leaf_ar, *leaf_br;
  const int *a_soff, *a_sidx, *a_list;
  const int *b_soff, *b_skey;
  int *item_a, *item_blo;
  u32* item_cnt;
  PairTest pt;
  const int *b_sidx, *b_list;
  DB_HD void operator()(long long i) const {
    const int leaf = upper_bound_n(leaf_ioff, nleaves + 1, (int)i) - 1;
    const int j = (int)i - leaf_ioff[leaf];
    const int ar = leaf_ar[leaf], br = leaf_br[leaf];
    const int a = a_sidx[a_soff[ar] + j];
    const int kcol = a_list[3 * (size_t)a + 1];
    const int* keys = b_skey + b_soff[br];
    const int len = b_soff[br + 1] - b_soff[br];
    const int lo = lower_bound_n(keys, len, kcol), hi = upper_bound_n(keys, len, kcol);
    item_a[i] = a;
    item_blo[i] = b_soff[br] + lo;
    u32 cnt = (u32)(hi - lo);
    if (pt.active()) {
      const int row = a_list[3 * (size_t)a];
      cnt = 0;
      for (int t = lo; t < hi; ++t) {
        const int b = b_sidx[b_soff[br] + t];
        cnt += pt.keep(a, b, row, b_list[3 * (size_t)b + 1]) ? 1u : 0u;
      }
    }
    item_cnt[i] = cnt;
  }
};

struct EmitProducts {
  const int *item_a, *item_blo;
  const u32* item_off;  // nitems + 1
  const int* b_sidx;
  int *prod_a, *prod_b;
  PairTest pt;
  const int *a_list, *b_list;
  DB_HD void operator()(long long i) const {
    const u32 off = item_off[i], cnt = item_off[i + 1] - off;
    const int a = item_a[i], blo = item_blo[i];
    if (!pt.active()) {
      for (u32 t = 0; t < cnt; ++t) {
        prod_a[off + t] = a;
        prod_b[off + t] = b_sidx[blo + (int)t];
      }
      return;
    }
    const int row = a_list[3 * (size_t)a];
    u32 w = 0;
    for (int t = 0; w < cnt; ++t) {  // the surviving pairs, in order
      const int b = b_sidx[blo + t];
      if (pt.keep(a, b, row, b_list[3 * (size_t)b + 1])) {
        prod_a[off + w] = a;
        prod_b[off + w] = b;
        ++w;
      }
    }
  }
};

// (row, col) -> slot; direct array when the block grid is small, else open addressing (the reference: one hash table per C row,
// src/utils/dbcsr_hash_table.f90 -- affects speed only)
struct CTable {
  u64* keys = nullptr;  // nullptr: direct
  u32* val = nullptr;
  u32 mask = 0;
  int ncols = 0;
  DB_HD u32 find_or_insert(int row, int col) const {
    if (keys == nullptr) return (u32)(row - 1) * (u32)ncols + (u32)(col - 1);
    const u64 key = ((u64)(u32)row << 32) | (u32)col;
    u32 h = (u32)((key * 0x9E3779B97F4A7C15ull) >> 32) & mask;
    for (;;) {
      const u64 old = atomic_cas_u64(&keys[h], 0ull, key);
      if (old == 0ull || old == key) return h;
      h = (h + 1) & mask;
    }
  }
  DB_HD u32 find(int row, int col) const {  // no insertion (nothing else inserts concurrently): kNoSlot when absent
    if (keys == nullptr) {
      const u32 s = (u32)(row - 1) * (u32)ncols + (u32)(col - 1);
      return val[s] == 0xFFFFFFFFu ? kNoSlot : s;
    }
    const u64 key = ((u64)(u32)row << 32) | (u32)col;
    u32 h = (u32)((key * 0x9E3779B97F4A7C15ull) >> 32) & mask;
    for (;;) {
      const u64 k = keys[h];
      if (k == key) return h;
      if (k == 0ull) return kNoSlot;
      h = (h + 1) & mask;
    }
  }
};

struct ReinsertBlocks {  // after the open-addressing table grew
  CTable t;
  const int *c_row, *c_col;
  DB_HD void operator()(long long id) const { t.val[t.find_or_insert(c_row[id], c_col[id])] = (u32)id + 1; }
};

struct InsertProducts {
  CTable t;
  const int *a_list, *b_list, *prod_a, *prod_b;
  u32* slot;
  int keep_sparsity;
  DB_HD void operator()(long long p) const {
    const int row = a_list[3 * (size_t)prod_a[p]], col = b_list[3 * (size_t)prod_b[p] + 1];
    if (keep_sparsity) {  // src/mm/dbcsr_mm_csr.F:307: only existing blocks receive products
      slot[p] = t.find(row, col);
      return;
    }
    const u32 s = t.find_or_insert(row, col);
    atomic_min_u32(&t.val[s], kNew | (u32)p);  // existing blocks hold their id (< kNew) and keep it
    slot[p] = s;
  }
};

struct FlagFirst {  // the product that touches a new block first creates it (src/mm/dbcsr_mm_csr.F:309-323)
  const u32 *val, *slot;
  const int *a_list, *b_list, *prod_a, *prod_b, *m_sizes, *n_sizes;
  u32* flag;
  u64* nze;
  DB_HD void operator()(long long p) const {
    const bool first = slot[p] != kNoSlot && val[slot[p]] == (kNew | (u32)p);
    flag[p] = first ? 1u : 0u;
    const int row = a_list[3 * (size_t)prod_a[p]], col = b_list[3 * (size_t)prod_b[p] + 1];
    nze[p] = first ? (u64)m_sizes[row - 1] * (u64)n_sizes[col - 1] : 0ull;
  }
};

struct WriteFirst {
  const u32 *flag, *fscan, *slot;
  const u64* zscan;
  const int *a_list, *b_list, *prod_a, *prod_b;
  int nblk0, ds0;
  u32* val;
  int *c_row, *c_col, *c_blkp;
  DB_HD void operator()(long long p) const {
    if (!flag[p]) return;
    const int id = nblk0 + (int)fscan[p] + 1;
    val[slot[p]] = (u32)id;
    c_row[id - 1] = a_list[3 * (size_t)prod_a[p]];
    c_col[id - 1] = b_list[3 * (size_t)prod_b[p] + 1];
    c_blkp[id - 1] = ds0 + (int)zscan[p] + 1;
  }
};

struct SizeMaps {
  const int *m_sizes, *n_sizes, *k_sizes;
  const int *m_map, *n_map, *k_map, *stack_map;
  int m_map_n, n_map_n, k_map_n, n_stacks;
};

struct BinProducts {  // stack number of every product (stack_map, src/mm/dbcsr_mm_csr.F:340-345)
  SizeMaps z;
  const int *a_list, *b_list, *prod_a, *prod_b;
  const u32* slot;
  u8* bin;
  u32* iota;
  DB_HD void operator()(long long p) const {
    const int a = prod_a[p], b = prod_b[p];
    const int m = z.m_sizes[a_list[3 * (size_t)a] - 1], k = z.k_sizes[a_list[3 * (size_t)a + 1] - 1], n = z.n_sizes[b_list[3 * (size_t)b + 1] - 1];
    int ws = kDropBin;
    if (m * n != 0 && k != 0 && slot[p] != kNoSlot) {
      const int w = z.n_stacks + 1;
      const int mm = m < z.m_map_n ? z.m_map[m] : w, mk = k < z.k_map_n ? z.k_map[k] : w, mn = n < z.n_map_n ? z.n_map[n] : w;
      ws = z.stack_map[((size_t)(mm - 1) * w + (mk - 1)) * w + (mn - 1)];
    }
    bin[p] = (u8)ws;
    iota[p] = (u32)p;
  }
};

struct BinStarts {
  const u8* bin_sorted;
  int n;
  int* bin_start;  // 257
  DB_HD void operator()(long long b) const { bin_start[b] = lower_bound_n(bin_sorted, n, (int)b); }
};

struct SliceInfo {  // traversal position at which every slice ends, datasize after it; totals
  const u32* item_off;
  const int* slice_end_item;
  const u32* fscan;
  const u64* zscan;
  int nslices, nprod, nblk0, ds0;
  u32* slice_end_pos;
  long long* info;  // [0] new blocks, [1] new elements, [2 + s] datasize after slice s
  DB_HD void operator()(long long s) const {
    if (s == nslices) {
      info[0] = (long long)fscan[nprod];
      info[1] = (long long)zscan[nprod];
      return;
    }
    const u32 e = item_off[slice_end_item[s]];
    slice_end_pos[s] = e;
    info[2 + s] = (long long)ds0 + (long long)zscan[e];
  }
};

// flush_stacks (src/mm/dbcsr_mm_csr.F:704-739) replayed on the per-stack lists of traversal positions: a stack that reaches
// mm_stack_size entries dispatches every stack above 3/4; a slice (multiply call) ends with a purge.
struct FlushSim {
  const int* bin_start;
  const u32* part_p;
  const u32* slice_end_pos;
  int nslices, nstacks, S, minfill, max_out;
  int *base, *fill;  // [256]
  u32* cand;         // [256]
  int* out;          // 4 per dispatch: ws, begin, size, slice
  int* nd;
  DB_HD void candidate(int b, u32 E) const {  // position of the product that fills stack b, if it comes before E
    const int cnt = bin_start[b + 1] - bin_start[b];
    const long long r = (long long)base[b] + S - 1;
    u32 c = kEmpty;
    if (r < cnt) {
      const u32 pos = part_p[(size_t)bin_start[b] + (size_t)r];
      if (pos < E) c = pos;
    }
    cand[b] = c;
  }
  DB_HD void count(int b, u32 x_excl) const {  // entries of stack b not yet dispatched with position < x_excl
    const int cnt = bin_start[b + 1] - bin_start[b];
    const int lo = base[b];
    const long long hi_l = (long long)lo + S;
    const int hi = hi_l < cnt ? (int)hi_l : cnt;
    fill[b] = hi > lo ? lower_bound_n(part_p + (size_t)bin_start[b] + lo, hi - lo, x_excl) : 0;
  }
  DB_HD void emit(int threshold, int slice) const {
    for (int i = 1; i <= nstacks; ++i) {
      if (fill[i] > threshold) {
        const int n = *nd;
        if (n < max_out) {
          out[4 * n] = i;
          out[4 * n + 1] = base[i];
          out[4 * n + 2] = fill[i];
          out[4 * n + 3] = slice;
        }
        *nd = n + 1;
        base[i] += fill[i];
      }
    }
  }
};

struct MakeKeys {  // (dispatch number, c_first) for stacks the accelerator driver sorts (stack_sort), else (dispatch number, rank)
  const long long* seq_start;  // nd + 1
  int nd;
  const int *d_ws, *d_begin;
  const u8* d_devord;
  const int* bin_start;
  const u32 *part_p, *slot, *val;
  const int* c_blkp;
  int lowbits;
  u64* key;
  u32* pout;
  DB_HD void operator()(long long i) const {
    const int d = upper_bound_n(seq_start, nd + 1, i) - 1;
    const u32 j = (u32)(i - seq_start[d]);
    const u32 p = part_p[(size_t)bin_start[d_ws[d]] + (size_t)d_begin[d] + j];
    const u32 low = d_devord[d] ? (u32)c_blkp[val[slot[p]] - 1] : j;
    key[i] = ((u64)d << lowbits) | low;
    pout[i] = p;
  }
};

// Tile order (not the reference's): inside a (slice, stack number) group the products are ordered by T x T tiles of C blocks, inside
// a tile by c_first -- every C block is then accumulated in ONE run of consecutive entries (all its products are adjacent) and a
// tile's A rows / B columns / C blocks stay L2 resident while it is worked on.
struct MakeKeysTiled {
  const long long* seq_start;  // nd + 1 (dispatch list of the flush rule: it partitions the same products)
  int nd;
  const int *d_ws, *d_begin, *d_group;
  const int* bin_start;
  const u32 *part_p, *slot, *val;
  const int *a_list, *b_list, *prod_a, *prod_b, *c_blkp;
  int lowbits, tilebits, tile, ntc;
  u64* key;
  u32* pout;
  DB_HD void operator()(long long i) const {
    const int d = upper_bound_n(seq_start, nd + 1, i) - 1;
    const u32 j = (u32)(i - seq_start[d]);
    const u32 p = part_p[(size_t)bin_start[d_ws[d]] + (size_t)d_begin[d] + j];
    const int row = a_list[3 * (size_t)prod_a[p]], col = b_list[3 * (size_t)prod_b[p] + 1];
    const u64 t = (u64)((row - 1) / tile) * (u64)ntc + (u64)((col - 1) / tile);
    key[i] = ((u64)d_group[d] << (lowbits + tilebits)) | (t << lowbits) | (u64)(u32)c_blkp[val[slot[p]] - 1];
    pout[i] = p;
  }
};

struct WriteStack3 {
  const u32 *psorted, *slot, *val;
  const int *a_list, *b_list, *prod_a, *prod_b, *c_blkp;
  int* out3;
  DB_HD void operator()(long long i) const {
    const u32 p = psorted[i];
    out3[3 * (size_t)i] = a_list[3 * (size_t)prod_a[p] + 2];
    out3[3 * (size_t)i + 1] = b_list[3 * (size_t)prod_b[p] + 2];
    out3[3 * (size_t)i + 2] = c_blkp[val[slot[p]] - 1];
  }
};

struct WriteParams7 {  // one dispatch entry in host order: products plist[offset ...]
  SizeMaps z;
  const u32 *plist, *slot, *val;
  const int *a_list, *b_list, *prod_a, *prod_b, *c_blkp;
  long long offset;
  int* out7;
  DB_HD void operator()(long long i) const {
    const u32 p = plist[(size_t)offset + (size_t)i];
    const int a = prod_a[p], b = prod_b[p];
    const int id = (int)val[slot[p]];
    int* o = out7 + 7 * (size_t)i;
    o[0] = z.m_sizes[a_list[3 * (size_t)a] - 1];
    o[1] = z.n_sizes[b_list[3 * (size_t)b + 1] - 1];
    o[2] = z.k_sizes[a_list[3 * (size_t)a + 1] - 1];
    o[3] = a_list[3 * (size_t)a + 2];
    o[4] = b_list[3 * (size_t)b + 2];
    o[5] = c_blkp[id - 1];
    o[6] = id;
  }
};

// ---------------------------------------------------------------------------------------------------------- executors
template <class F>
__global__ void for_each_kernel(long long n, F f) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) f(i);
}

__global__ void flushsim_kernel(FlushSim fs) {
  __shared__ u32 s_pe;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int b = tid; b < 256; b += nt) {
    fs.base[b] = 0;
    fs.fill[b] = 0;
  }
  if (tid == 0) *fs.nd = 0;
  __syncthreads();
  for (int s = 0; s < fs.nslices; ++s) {
    const u32 E = fs.slice_end_pos[s];
    for (;;) {
      for (int b = 1 + tid; b <= fs.nstacks; b += nt) fs.candidate(b, E);
      __syncthreads();
      if (tid == 0) {
        u32 pe = kEmpty;
        for (int b = 1; b <= fs.nstacks; ++b) pe = fs.cand[b] < pe ? fs.cand[b] : pe;
        s_pe = pe;
      }
      __syncthreads();
      const u32 pe = s_pe;
      const u32 x = pe == kEmpty ? E : pe + 1;
      for (int b = 1 + tid; b <= fs.nstacks; b += nt) fs.count(b, x);
      __syncthreads();
      if (tid == 0) fs.emit(pe == kEmpty ? 0 : fs.minfill, s);
      __syncthreads();
      if (pe == kEmpty) break;
    }
  }
}

struct DeviceExec {
  cudaStream_t st = nullptr;
  void* temp = nullptr;
  size_t temp_cap = 0;
  static constexpr bool kDevice = true;
  ~DeviceExec() {
    if (temp != nullptr) cudaFree(temp);
  }
  void* alloc(size_t n) {
    void* p = nullptr;
    return cudaMalloc(&p, n ? n : 1) == cudaSuccess ? p : nullptr;
  }
  void release(void* p) {
    if (p != nullptr) cudaFree(p);
  }
  int h2d(void* d, const void* h, size_t n) { return n == 0 || cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, st) == cudaSuccess ? 0 : -61; }
  int d2h(void* h, const void* d, size_t n) { return n == 0 || cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, st) == cudaSuccess ? 0 : -61; }
  int d2d(void* dst, const void* src, size_t n) { return n == 0 || cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, st) == cudaSuccess ? 0 : -61; }
  int fill(void* d, int byte, size_t n) { return n == 0 || cudaMemsetAsync(d, byte, n, st) == cudaSuccess ? 0 : -61; }
  int sync() { return cudaStreamSynchronize(st) == cudaSuccess ? 0 : -62; }
  template <class F>
  int for_each(long long n, const F& f) {
    if (n <= 0) return 0;
    const int threads = 256;
    long long blocks = (n + threads - 1) / threads;
    if (blocks > 148 * 32) blocks = 148 * 32;
    for_each_kernel<<<(unsigned)blocks, threads, 0, st>>>(n, f);
    return cudaGetLastError() == cudaSuccess ? 0 : -63;
  }
  int need_temp(size_t bytes) {
    if (bytes <= temp_cap) return 0;
    if (temp != nullptr) {
      cudaStreamSynchronize(st);
      cudaFree(temp);
    }
    temp = nullptr;
    temp_cap = 0;
    if (cudaMalloc(&temp, bytes + bytes / 4) != cudaSuccess) return -64;
    temp_cap = bytes + bytes / 4;
    return 0;
  }
  template <class T>
  int scan(const T* in, T* out, long long n) {  // exclusive sum
    if (n <= 0) return 0;
    size_t bytes = 0;
    if (cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, st) != cudaSuccess) return -65;
    if (need_temp(bytes) != 0) return -64;
    return cub::DeviceScan::ExclusiveSum(temp, bytes, in, out, (int)n, st) == cudaSuccess ? 0 : -65;
  }
  template <class K>
  int sort_pairs(const K* kin, K* kout, const u32* vin, u32* vout, long long n, int end_bit) {  // stable, ascending
    if (n <= 0) return 0;
    size_t bytes = 0;
    if (cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, (int)n, 0, end_bit, st) != cudaSuccess) return -66;
    if (need_temp(bytes) != 0) return -64;
    return cub::DeviceRadixSort::SortPairs(temp, bytes, kin, kout, vin, vout, (int)n, 0, end_bit, st) == cudaSuccess ? 0 : -66;
  }
  int flushsim(const FlushSim& fs) {
    flushsim_kernel<<<1, 256, 0, st>>>(fs);
    return cudaGetLastError() == cudaSuccess ? 0 : -63;
  }
};

struct HostExec {
  static constexpr bool kDevice = false;
  void* alloc(size_t n) { return std::malloc(n ? n : 1); }
  void release(void* p) { std::free(p); }
  int h2d(void* d, const void* h, size_t n) {
    if (n) std::memcpy(d, h, n);
    return 0;
  }
  int d2h(void* h, const void* d, size_t n) { return h2d(h, d, n); }
  int d2d(void* dst, const void* src, size_t n) { return h2d(dst, src, n); }
  int fill(void* d, int byte, size_t n) {
    if (n) std::memset(d, byte, n);
    return 0;
  }
  int sync() { return 0; }
  template <class F>
  int for_each(long long n, const F& f) {
    for (long long i = 0; i < n; ++i) f(i);
    return 0;
  }
  template <class T>
  int scan(const T* in, T* out, long long n) {
    T run = 0;
    for (long long i = 0; i < n; ++i) {
      const T v = in[i];
      out[i] = run;
      run += v;
    }
    return 0;
  }
  template <class K>
  int sort_pairs(const K* kin, K* kout, const u32* vin, u32* vout, long long n, int end_bit) {
    std::vector<long long> order((size_t)n);
    std::iota(order.begin(), order.end(), 0ll);
    const u64 keep = end_bit >= 64 ? ~0ull : ((1ull << end_bit) - 1ull);
    std::stable_sort(order.begin(), order.end(), [&](long long x, long long y) { return ((u64)kin[x] & keep) < ((u64)kin[y] & keep); });
    for (long long i = 0; i < n; ++i) {
      kout[i] = kin[order[(size_t)i]];
      vout[i] = vin[order[(size_t)i]];
    }
    return 0;
  }
  int flushsim(const FlushSim& fs) {  // the kernel's phases, one "thread" at a time
    for (int b = 0; b < 256; ++b) fs.base[b] = fs.fill[b] = 0;
    *fs.nd = 0;
    for (int s = 0; s < fs.nslices; ++s) {
      const u32 E = fs.slice_end_pos[s];
      for (;;) {
        u32 pe = kEmpty;
        for (int b = 1; b <= fs.nstacks; ++b) {
          fs.candidate(b, E);
          pe = std::min(pe, fs.cand[b]);
        }
        const u32 x = pe == kEmpty ? E : pe + 1;
        for (int b = 1; b <= fs.nstacks; ++b) fs.count(b, x);
        fs.emit(pe == kEmpty ? 0 : fs.minfill, s);
        if (pe == kEmpty) break;
      }
    }
    return 0;
  }
};

// ---------------------------------------------------------------------------------------------------------- builder
template <class X>
class Builder final : public IDeviceBuilder {
  struct Buf {
    void* p = nullptr;
    size_t cap = 0;
  };
  X x_;
  // persistent over the ticks of a multiply
  Buf c_row_, c_col_, c_blkp_, tkeys_, tval_;
  CTable table_;
  long long table_slots_ = 0;
  bool table_dense_ = false;
  int nblk_ = 0, datasize_ = 0;
  // per tick
  Buf a_list_, b_list_, ints_, item_a_, item_blo_, item_cnt_, item_off_, prod_a_, prod_b_, slot_, w32a_, w32b_, w64a_, w64b_, bin_, bin_s_, iota_,
      part_p_, sim_, disp_, out3_, p7_, info_, sa_idx_, sa_key_, sb_idx_, sb_key_, dtab_, ptest_;
  SizeMaps maps_{};
  const int* d_bin_start_ = nullptr;
  long long nprod_ = 0, nkept_ = 0;
  std::vector<int> h_bin_start_;
  int tile_ = 0;             // > 0: tile order (set_tile_order)
  bool tiled_last_ = false;  // the last build used it

  int ensure(Buf& b, size_t bytes, bool keep = false, size_t keep_bytes = 0) {
    if (bytes <= b.cap) return 0;
    size_t cap = bytes + bytes / 4 + 256;
    void* q = x_.alloc(cap);
    if (q == nullptr) {
      cap = bytes;
      q = x_.alloc(cap);
      if (q == nullptr) return -40;
    }
    if (keep && b.p != nullptr && keep_bytes > 0 && x_.d2d(q, b.p, keep_bytes) != 0) return -61;
    if (b.p != nullptr) {
      if (X::kDevice) x_.sync();  // everything that still reads the old buffer is on this stream
      x_.release(b.p);
    }
    b.p = q;
    b.cap = cap;
    return 0;
  }
  template <class T>
  static T* ptr(const Buf& b) {
    return static_cast<T*>(b.p);
  }

 public:
  explicit Builder(void* stream) { set_stream(x_, stream); }
  static void set_stream(DeviceExec& x, void* s) { x.st = static_cast<cudaStream_t>(s); }
  static void set_stream(HostExec&, void*) {}
  ~Builder() override {
    x_.sync();
    for (Buf* b : {&c_row_, &c_col_, &c_blkp_, &tkeys_, &tval_, &a_list_, &b_list_, &ints_, &item_a_, &item_blo_, &item_cnt_, &item_off_, &prod_a_,
                   &prod_b_, &slot_, &w32a_, &w32b_, &w64a_, &w64b_, &bin_, &bin_s_, &iota_, &part_p_, &sim_, &disp_, &out3_, &p7_, &info_, &sa_idx_,
                   &sa_key_, &sb_idx_, &sb_key_, &dtab_, &ptest_})
      if (b->p != nullptr) x_.release(b->p);
  }

  int reset() override {
    nblk_ = 0;
    datasize_ = 0;
    if (table_slots_ > 0) {
      if (x_.fill(tval_.p, 0xFF, sizeof(u32) * (size_t)table_slots_) != 0) return -61;
      if (!table_dense_ && x_.fill(tkeys_.p, 0, sizeof(u64) * (size_t)table_slots_) != 0) return -61;
    }
    return 0;
  }

  // the C table must be able to take `extra` more blocks
  int prepare_table(long long nrows, long long ncols, long long extra) {
    if (nrows * ncols <= dense_limit()) {
      const long long slots = std::max(1ll, nrows * ncols);
      table_.ncols = (int)ncols;
      if (table_slots_ == slots && table_dense_) return 0;
      if (nblk_ != 0) return -67;  // the grid cannot change inside a multiply
      if (int rc = ensure(tval_, sizeof(u32) * (size_t)slots)) return rc;
      if (x_.fill(tval_.p, 0xFF, sizeof(u32) * (size_t)slots) != 0) return -61;
      table_slots_ = slots;
      table_dense_ = true;
      table_.keys = nullptr;
      table_.val = ptr<u32>(tval_);
      table_.mask = 0;
      table_.ncols = (int)ncols;
      return 0;
    }
    const long long need = 2 * ((long long)nblk_ + extra) + 16;
    if (!table_dense_ && table_slots_ >= need) return 0;
    long long slots = 1ll << 16;
    while (slots < need) slots <<= 1;
    if (slots > (1ll << 31)) return -68;
    // grow: fresh arrays, existing blocks re-inserted
    Buf nk, nv;
    if (int rc = ensure(nk, sizeof(u64) * (size_t)slots)) return rc;
    if (int rc = ensure(nv, sizeof(u32) * (size_t)slots)) return rc;
    if (x_.fill(nk.p, 0, sizeof(u64) * (size_t)slots) != 0 || x_.fill(nv.p, 0xFF, sizeof(u32) * (size_t)slots) != 0) return -61;
    if (X::kDevice) x_.sync();
    if (tkeys_.p != nullptr) x_.release(tkeys_.p);
    if (tval_.p != nullptr) x_.release(tval_.p);
    tkeys_ = nk;
    tval_ = nv;
    table_slots_ = slots;
    table_dense_ = false;
    table_.keys = ptr<u64>(tkeys_);
    table_.val = ptr<u32>(tval_);
    table_.mask = (u32)(slots - 1);
    table_.ncols = (int)ncols;
    if (nblk_ > 0) {
      ReinsertBlocks f{table_, ptr<int>(c_row_), ptr<int>(c_col_)};
      if (int rc = x_.for_each(nblk_, f)) return rc;
    }
    return 0;
  }

  // existing C blocks: ids 1..nblks in list order at the given offsets
  int preset(int nrows, int ncols, const int* rows, const int* cols, const int* blk_p, int nblks, int datasize) override {
    if (nblk_ != 0 || nblks < 0) return -2;
    if (int rc = prepare_table(nrows, ncols, nblks)) return rc;
    const size_t bytes = sizeof(int) * (size_t)std::max(nblks, 1);
    if (int rc = ensure(c_row_, bytes)) return rc;
    if (int rc = ensure(c_col_, bytes)) return rc;
    if (int rc = ensure(c_blkp_, bytes)) return rc;
    if (nblks > 0) {
      if (x_.h2d(c_row_.p, rows, sizeof(int) * (size_t)nblks) != 0 || x_.h2d(c_col_.p, cols, sizeof(int) * (size_t)nblks) != 0 ||
          x_.h2d(c_blkp_.p, blk_p, sizeof(int) * (size_t)nblks) != 0)
        return -61;
      ReinsertBlocks f{table_, ptr<int>(c_row_), ptr<int>(c_col_)};
      if (int rc = x_.for_each(nblks, f)) return rc;
    }
    nblk_ = nblks;
    datasize_ = datasize;
    return 0;
  }

  int build(LocalMultiply& mm, const Idx3* a_sorted, int na, const std::vector<std::pair<int, int>>& slices, const Idx3* b_sorted, int nb,
            DevBuildResult& out, const DevBuildOptions& opt) override {
    const Config& cfg = mm.config();
    const int nstacks = mm.nstacks();
    out = DevBuildResult();
    out.nblk_before = out.nblk_after = nblk_;
    out.datasize_before = out.datasize_after = datasize_;
    out.slice_datasize.assign(slices.size(), datasize_);
    nprod_ = nkept_ = 0;
    if (nstacks > kMaxBins) return -60;
    const int nslices = (int)slices.size();
    if (nslices == 0) return 0;

    // ---- host: the recursion of sparse_multrec, leaves only; distinct list ranges
    std::vector<LocalMultiply::Leaf> leaves;
    std::vector<int> slice_end_leaf((size_t)nslices);
    for (int s = 0; s < nslices; ++s) {
      if (slices[(size_t)s].second >= slices[(size_t)s].first && nb > 0)
        mm.plan(a_sorted, slices[(size_t)s].first, slices[(size_t)s].second, b_sorted, nb, leaves);
      slice_end_leaf[(size_t)s] = (int)leaves.size();
    }
    const int nleaves = (int)leaves.size();
    std::map<std::pair<int, int>, int> a_ids, b_ids;
    std::vector<int> a_start, a_soff{0}, b_start, b_soff{0}, leaf_ar((size_t)nleaves), leaf_br((size_t)nleaves), leaf_ioff((size_t)nleaves + 1, 0);
    for (int l = 0; l < nleaves; ++l) {
      const auto& lf = leaves[(size_t)l];
      auto ia = a_ids.find({lf.ai, lf.af});
      if (ia == a_ids.end()) {
        ia = a_ids.emplace(std::make_pair(lf.ai, lf.af), (int)a_start.size()).first;
        a_start.push_back(lf.ai - 1);
        a_soff.push_back(a_soff.back() + (lf.af - lf.ai + 1));
      }
      auto ib = b_ids.find({lf.bi, lf.bf});
      if (ib == b_ids.end()) {
        ib = b_ids.emplace(std::make_pair(lf.bi, lf.bf), (int)b_start.size()).first;
        b_start.push_back(lf.bi - 1);
        b_soff.push_back(b_soff.back() + (lf.bf - lf.bi + 1));
      }
      leaf_ar[(size_t)l] = ia->second;
      leaf_br[(size_t)l] = ib->second;
      const long long next = (long long)leaf_ioff[(size_t)l] + (lf.af - lf.ai + 1);
      if (next > 0x7fffffffll) return -69;
      leaf_ioff[(size_t)l + 1] = (int)next;
    }
    const int nitems = leaf_ioff[(size_t)nleaves];
    std::vector<int> slice_end_item((size_t)nslices);
    for (int s = 0; s < nslices; ++s) slice_end_item[(size_t)s] = leaf_ioff[(size_t)slice_end_leaf[(size_t)s]];
    const int nar = (int)a_start.size(), nbr = (int)b_start.size();

    // ---- uploads: lists, plan tables, size maps (k sizes change from tick to tick)
    if (int rc = ensure(a_list_, sizeof(int) * 3 * (size_t)std::max(na, 1))) return rc;
    if (int rc = ensure(b_list_, sizeof(int) * 3 * (size_t)std::max(nb, 1))) return rc;
    if (x_.h2d(a_list_.p, a_sorted, sizeof(int) * 3 * (size_t)na) != 0 || x_.h2d(b_list_.p, b_sorted, sizeof(int) * 3 * (size_t)nb) != 0) return -61;
    const int* d_a_list = ptr<int>(a_list_);
    const int* d_b_list = ptr<int>(b_list_);
    std::vector<int> pack;
    auto push = [&pack](const std::vector<int>& v) {
      const size_t off = pack.size();
      pack.insert(pack.end(), v.begin(), v.end());
      return off;
    };
    const size_t o_as = push(a_start), o_aso = push(a_soff), o_bs = push(b_start), o_bso = push(b_soff), o_lar = push(leaf_ar), o_lbr = push(leaf_br),
                 o_lio = push(leaf_ioff), o_sei = push(slice_end_item), o_ms = push(mm.m_sizes()), o_ns = push(mm.n_sizes()), o_ks = push(mm.k_sizes()),
                 o_mm = push(mm.m_map()), o_nm = push(mm.n_map()), o_km = push(mm.k_map()), o_sm = push(mm.stack_map());
    if (int rc = ensure(ints_, sizeof(int) * pack.size())) return rc;
    if (x_.h2d(ints_.p, pack.data(), sizeof(int) * pack.size()) != 0) return -61;
    // (`pack` and the lists are pageable host memory: cudaMemcpyAsync returns once they have been staged)
    const int* I = ptr<int>(ints_);
    maps_.m_sizes = I + o_ms, maps_.n_sizes = I + o_ns, maps_.k_sizes = I + o_ks;
    maps_.m_map = I + o_mm, maps_.n_map = I + o_nm, maps_.k_map = I + o_km, maps_.stack_map = I + o_sm;
    maps_.m_map_n = (int)mm.m_map().size(), maps_.n_map_n = (int)mm.n_map().size(), maps_.k_map_n = (int)mm.k_map().size();
    maps_.n_stacks = cfg.n_stacks;
    // ---- the pair tests of this tick (on-the-fly filter, symmetric-product skipping)
    PairTest pt;
    {
      const bool filt = opt.a_norms != nullptr && opt.b_norms != nullptr && opt.row_eps != nullptr;
      const size_t nrows_l = mm.m_sizes().size(), ncols_l = mm.n_sizes().size();
      const size_t nf = filt ? (size_t)na + (size_t)nb + nrows_l : 0;
      const size_t ng = opt.c_sym ? (opt.global_rows != nullptr ? nrows_l : 0) + (opt.global_cols != nullptr ? ncols_l : 0) : 0;
      if (nf + ng > 0) {
        if (int rc = ensure(ptest_, sizeof(float) * nf + sizeof(int) * ng + 16)) return rc;
        float* f = ptr<float>(ptest_);
        if (filt) {
          if (x_.h2d(f, opt.a_norms, sizeof(float) * (size_t)na) != 0 || x_.h2d(f + na, opt.b_norms, sizeof(float) * (size_t)nb) != 0 ||
              x_.h2d(f + na + nb, opt.row_eps, sizeof(float) * nrows_l) != 0)
            return -61;
          pt.a_norms = f, pt.b_norms = f + na, pt.row_eps = f + na + nb;
        }
        int* g = reinterpret_cast<int*>(f + nf);
        if (opt.c_sym && opt.global_rows != nullptr) {
          if (x_.h2d(g, opt.global_rows, sizeof(int) * nrows_l) != 0) return -61;
          pt.grow = g;
          g += nrows_l;
        }
        if (opt.c_sym && opt.global_cols != nullptr) {
          if (x_.h2d(g, opt.global_cols, sizeof(int) * ncols_l) != 0) return -61;
          pt.gcol = g;
        }
      }
      pt.c_sym = opt.c_sym ? 1 : 0;
    }

    if (nitems > 0) {
      // ---- CSR order of every distinct range
      const int ta = a_soff.back(), tb = b_soff.back();
      if (int rc = ensure(sa_idx_, sizeof(int) * (size_t)ta)) return rc;
      if (int rc = ensure(sa_key_, sizeof(int) * (size_t)ta)) return rc;
      if (int rc = ensure(sb_idx_, sizeof(int) * (size_t)tb)) return rc;
      if (int rc = ensure(sb_key_, sizeof(int) * (size_t)tb)) return rc;
      if (int rc = x_.for_each(ta, RankRange{d_a_list, I + o_as, I + o_aso, nar, ptr<int>(sa_idx_), ptr<int>(sa_key_)})) return rc;
      if (int rc = x_.for_each(tb, RankRange{d_b_list, I + o_bs, I + o_bso, nbr, ptr<int>(sb_idx_), ptr<int>(sb_key_)})) return rc;
      // ---- products per item, traversal positions
      if (int rc = ensure(item_a_, sizeof(int) * (size_t)nitems)) return rc;
      if (int rc = ensure(item_blo_, sizeof(int) * (size_t)nitems)) return rc;
      if (int rc = ensure(item_cnt_, sizeof(u32) * ((size_t)nitems + 1))) return rc;
      if (int rc = ensure(item_off_, sizeof(u32) * ((size_t)nitems + 1))) return rc;
      if (x_.fill(ptr<u32>(item_cnt_) + nitems, 0, sizeof(u32)) != 0) return -61;
      CountItems ci{I + o_lio, nleaves, I + o_lar, I + o_lbr, I + o_aso, ptr<int>(sa_idx_), d_a_list, I + o_bso, ptr<int>(sb_key_),
                    ptr<int>(item_a_), ptr<int>(item_blo_), ptr<u32>(item_cnt_), pt, ptr<int>(sb_idx_), d_b_list};
      if (int rc = x_.for_each(nitems, ci)) return rc;
      if (int rc = x_.scan(ptr<u32>(item_cnt_), ptr<u32>(item_off_), (long long)nitems + 1)) return rc;
      u32 total = 0;
      if (x_.d2h(&total, ptr<u32>(item_off_) + nitems, sizeof(u32)) != 0 || x_.sync() != 0) return -62;  // message 1: product count
      if (total >= kNew) return -69;
      nprod_ = total;
    }
    else {
      if (int rc = ensure(item_off_, sizeof(u32))) return rc;
      if (x_.fill(item_off_.p, 0, sizeof(u32)) != 0) return -61;
    }
    out.nprod = nprod_;
    const long long N = nprod_;
    const long long nrows = (long long)mm.m_sizes().size(), ncols = (long long)mm.n_sizes().size();
    if (int rc = prepare_table(nrows, ncols, N)) return rc;
    {
      long long cap = (long long)nblk_ + N;
      if (table_dense_) cap = std::min(cap, nrows * ncols);
      cap = std::max(cap, 1ll);
      if (int rc = ensure(c_row_, sizeof(int) * (size_t)cap, true, sizeof(int) * (size_t)nblk_)) return rc;
      if (int rc = ensure(c_col_, sizeof(int) * (size_t)cap, true, sizeof(int) * (size_t)nblk_)) return rc;
      if (int rc = ensure(c_blkp_, sizeof(int) * (size_t)cap, true, sizeof(int) * (size_t)nblk_)) return rc;
    }
    if (int rc = ensure(prod_a_, sizeof(int) * (size_t)std::max(N, 1ll))) return rc;
    if (int rc = ensure(prod_b_, sizeof(int) * (size_t)std::max(N, 1ll))) return rc;
    if (int rc = ensure(slot_, sizeof(u32) * (size_t)std::max(N, 1ll))) return rc;
    if (int rc = ensure(w32a_, sizeof(u32) * ((size_t)N + 1))) return rc;
    if (int rc = ensure(w32b_, sizeof(u32) * ((size_t)N + 1))) return rc;
    if (int rc = ensure(w64a_, sizeof(u64) * ((size_t)N + 1))) return rc;
    if (int rc = ensure(w64b_, sizeof(u64) * ((size_t)N + 1))) return rc;
    if (int rc = ensure(bin_, (size_t)N + 1)) return rc;
    if (int rc = ensure(bin_s_, (size_t)N + 1)) return rc;
    if (int rc = ensure(iota_, sizeof(u32) * (size_t)std::max(N, 1ll))) return rc;
    if (int rc = ensure(part_p_, sizeof(u32) * (size_t)std::max(N, 1ll))) return rc;
    int* d_prod_a = ptr<int>(prod_a_);
    int* d_prod_b = ptr<int>(prod_b_);
    u32* d_slot = ptr<u32>(slot_);
    u32 *d_flag = ptr<u32>(w32a_), *d_fscan = ptr<u32>(w32b_);
    u64 *d_nze = ptr<u64>(w64a_), *d_zscan = ptr<u64>(w64b_);
    if (N > 0) {
      EmitProducts ep{ptr<int>(item_a_), ptr<int>(item_blo_), ptr<u32>(item_off_), ptr<int>(sb_idx_), d_prod_a, d_prod_b, pt, d_a_list, d_b_list};
      if (int rc = x_.for_each(nitems, ep)) return rc;
      // ---- C blocks: earliest toucher per (row, col), block numbers and offsets by scans
      if (int rc = x_.for_each(N, InsertProducts{table_, d_a_list, d_b_list, d_prod_a, d_prod_b, d_slot, opt.keep_sparsity ? 1 : 0})) return rc;
      if (int rc = x_.for_each(N, FlagFirst{table_.val, d_slot, d_a_list, d_b_list, d_prod_a, d_prod_b, maps_.m_sizes, maps_.n_sizes, d_flag, d_nze}))
        return rc;
    }
    if (x_.fill(d_flag + N, 0, sizeof(u32)) != 0 || x_.fill(d_nze + N, 0, sizeof(u64)) != 0) return -61;
    if (int rc = x_.scan(d_flag, d_fscan, N + 1)) return rc;
    if (int rc = x_.scan(d_nze, d_zscan, N + 1)) return rc;
    if (N > 0) {
      WriteFirst wf{d_flag, d_fscan, d_slot, d_zscan, d_a_list, d_b_list, d_prod_a, d_prod_b, nblk_, datasize_, table_.val,
                    ptr<int>(c_row_), ptr<int>(c_col_), ptr<int>(c_blkp_)};
      if (int rc = x_.for_each(N, wf)) return rc;
      // ---- stack number per product, stable partition
      if (int rc = x_.for_each(N, BinProducts{maps_, d_a_list, d_b_list, d_prod_a, d_prod_b, d_slot, ptr<u8>(bin_), ptr<u32>(iota_)})) return rc;
      if (int rc = x_.sort_pairs(ptr<u8>(bin_), ptr<u8>(bin_s_), ptr<u32>(iota_), ptr<u32>(part_p_), N, 8)) return rc;
    }
    // ---- flush rule; slice ends; totals
    const long long max_out_l = N / std::max(1, cfg.mm_stack_size * 3 / 4 + 1) + (long long)nslices * nstacks + 16;
    if (max_out_l > 0x7ffffff) return -69;
    const int max_out = (int)max_out_l;
    // sim_: bin_start[257] | base[256] | fill[256] | cand[256] | nd | slice_end_pos[nslices]
    if (int rc = ensure(sim_, sizeof(int) * (size_t)(257 + 3 * 256 + 1 + nslices))) return rc;
    if (int rc = ensure(disp_, sizeof(int) * 4 * (size_t)max_out)) return rc;
    if (int rc = ensure(info_, sizeof(long long) * (size_t)(2 + nslices))) return rc;
    int* d_bin_start = ptr<int>(sim_);
    d_bin_start_ = d_bin_start;
    u32* d_slice_end = reinterpret_cast<u32*>(d_bin_start + 257 + 3 * 256 + 1);
    if (int rc = x_.for_each(257, BinStarts{ptr<u8>(bin_s_), (int)N, d_bin_start})) return rc;
    SliceInfo si{ptr<u32>(item_off_), I + o_sei, d_fscan, d_zscan, nslices, (int)N, nblk_, datasize_, d_slice_end, ptr<long long>(info_)};
    if (int rc = x_.for_each(nslices + 1, si)) return rc;
    FlushSim fs{d_bin_start, ptr<u32>(part_p_), d_slice_end, nslices, nstacks, cfg.mm_stack_size, cfg.mm_stack_size * 3 / 4, max_out,
                d_bin_start + 257, d_bin_start + 257 + 256, reinterpret_cast<u32*>(d_bin_start + 257 + 512), ptr<int>(disp_),
                d_bin_start + 257 + 768};
    if (int rc = x_.flushsim(fs)) return rc;
    // ---- message 2: totals, slice datasizes, dispatch list
    std::vector<long long> info((size_t)(2 + nslices));
    int nd = 0;
    if (x_.d2h(info.data(), info_.p, sizeof(long long) * info.size()) != 0 || x_.d2h(&nd, d_bin_start + 257 + 768, sizeof(int)) != 0 || x_.sync() != 0)
      return -62;
    if (nd > max_out) return -69;
    h_bin_start_.assign(257, 0);
    if (x_.d2h(h_bin_start_.data(), d_bin_start, sizeof(int) * 257) != 0 || x_.sync() != 0) return -62;
    if ((long long)datasize_ + info[1] > 0x7fffffffll) return -42;  // offsets are int32
    std::vector<int> disp(4 * (size_t)nd);
    if (nd > 0 && (x_.d2h(disp.data(), disp_.p, sizeof(int) * disp.size()) != 0 || x_.sync() != 0)) return -62;
    out.nblk_after = nblk_ + (int)info[0];
    out.datasize_after = datasize_ + (int)info[1];
    for (int s = 0; s < nslices; ++s) out.slice_datasize[(size_t)s] = (int)info[(size_t)(2 + s)];
    out.dispatch.resize((size_t)nd);
    std::vector<long long> seq((size_t)nd + 1, 0);
    std::vector<int> d_ws((size_t)nd), d_begin((size_t)nd);
    std::vector<u8> d_devord((size_t)nd);
    for (int d = 0; d < nd; ++d) {
      DevDispatch& e = out.dispatch[(size_t)d];
      e.ws = disp[4 * (size_t)d];
      e.begin = disp[4 * (size_t)d + 1];
      e.size = disp[4 * (size_t)d + 2];
      e.slice = disp[4 * (size_t)d + 3];
      e.seq_start = seq[(size_t)d];
      seq[(size_t)d + 1] = seq[(size_t)d] + e.size;
      d_ws[(size_t)d] = e.ws;
      d_begin[(size_t)d] = e.begin;
      d_devord[(size_t)d] = (cfg.stack_sort && device_ordered(cfg, mm.descr(e.ws))) ? 1 : 0;
    }
    nkept_ = seq[(size_t)nd];
    nblk_ = out.nblk_after;
    datasize_ = out.datasize_after;
    tiled_last_ = false;
    if (nkept_ == 0) return 0;
    bool all_devord = cfg.stack_sort != 0;
    for (int d = 0; d < nd; ++d) all_devord = all_devord && d_devord[(size_t)d] != 0;
    if (tile_ > 0 && all_devord) {
      const int rc_t = build_tiled(mm, nslices, nd, seq, d_ws, d_begin, out);
      if (rc_t != -70) return rc_t;  // -70: keys would not fit => reference order below
    }
    // ---- device order of every stack: one stable sort over (dispatch number, c_first | rank)
    int lowbits = 1, dbits = 1;
    while ((1ll << lowbits) <= std::max((long long)datasize_, (long long)cfg.mm_stack_size)) ++lowbits;
    while ((1ll << dbits) < nd) ++dbits;
    const size_t tab_bytes = sizeof(long long) * ((size_t)nd + 1) + sizeof(int) * 2 * (size_t)nd + (size_t)nd;
    if (int rc = ensure(dtab_, tab_bytes + 64)) return rc;
    long long* t_seq = ptr<long long>(dtab_);
    int* t_ws = reinterpret_cast<int*>(t_seq + nd + 1);
    int* t_begin = t_ws + nd;
    u8* t_devord = reinterpret_cast<u8*>(t_begin + nd);
    if (x_.h2d(t_seq, seq.data(), sizeof(long long) * ((size_t)nd + 1)) != 0 || x_.h2d(t_ws, d_ws.data(), sizeof(int) * (size_t)nd) != 0 ||
        x_.h2d(t_begin, d_begin.data(), sizeof(int) * (size_t)nd) != 0 || x_.h2d(t_devord, d_devord.data(), (size_t)nd) != 0)
      return -61;
    u64 *d_key = d_nze, *d_key_s = d_zscan;  // the scan arrays are dead from here on
    u32 *d_p = d_flag, *d_p_s = d_fscan;
    MakeKeys mk{t_seq, nd, t_ws, t_begin, t_devord, d_bin_start, ptr<u32>(part_p_), d_slot, table_.val, ptr<int>(c_blkp_), lowbits, d_key, d_p};
    if (int rc = x_.for_each(nkept_, mk)) return rc;
    if (int rc = x_.sort_pairs(d_key, d_key_s, d_p, d_p_s, nkept_, lowbits + dbits)) return rc;
    if (int rc = ensure(out3_, sizeof(int) * 3 * (size_t)nkept_)) return rc;
    WriteStack3 ws3{d_p_s, d_slot, table_.val, d_a_list, d_b_list, d_prod_a, d_prod_b, ptr<int>(c_blkp_), ptr<int>(out3_)};
    if (int rc = x_.for_each(nkept_, ws3)) return rc;
    return 0;
  }

  void set_tile_order(int tile) override { tile_ = tile > 0 ? tile : 0; }

  // stacks of the tile order: groups (slice, stack number) in dispatch order, each cut into stacks of mm_stack_size entries
  int build_tiled(LocalMultiply& mm, int nslices, int nd, const std::vector<long long>& seq, const std::vector<int>& d_ws, const std::vector<int>& d_begin,
                  DevBuildResult& out) {
    const Config& cfg = mm.config();
    const int nstacks = mm.nstacks();
    const long long nrows = (long long)mm.m_sizes().size(), ncols = (long long)mm.n_sizes().size();
    const long long ntr = (nrows + tile_ - 1) / tile_, ntc = (ncols + tile_ - 1) / tile_;
    std::vector<long long> gcount((size_t)nslices * (nstacks + 1), 0);
    std::vector<int> d_group((size_t)nd);
    for (int d = 0; d < nd; ++d) {
      const int gi = out.dispatch[(size_t)d].slice * (nstacks + 1) + out.dispatch[(size_t)d].ws;
      gcount[(size_t)gi] += out.dispatch[(size_t)d].size;
      d_group[(size_t)d] = gi;
    }
    int lowbits = 1, tilebits = 1, gbits = 1;
    while ((1ll << lowbits) <= (long long)datasize_) ++lowbits;
    while ((1ll << tilebits) < ntr * ntc) ++tilebits;
    while ((1ll << gbits) < (long long)gcount.size()) ++gbits;
    if (lowbits + tilebits + gbits > 63 || ntc > 0x7fffffffll) return -70;
    const size_t tab_bytes = sizeof(long long) * ((size_t)nd + 1) + sizeof(int) * 3 * (size_t)nd;
    if (int rc = ensure(dtab_, tab_bytes + 64)) return rc;
    long long* t_seq = ptr<long long>(dtab_);
    int* t_ws = reinterpret_cast<int*>(t_seq + nd + 1);
    int* t_begin = t_ws + nd;
    int* t_group = t_begin + nd;
    if (x_.h2d(t_seq, seq.data(), sizeof(long long) * ((size_t)nd + 1)) != 0 || x_.h2d(t_ws, d_ws.data(), sizeof(int) * (size_t)nd) != 0 ||
        x_.h2d(t_begin, d_begin.data(), sizeof(int) * (size_t)nd) != 0 || x_.h2d(t_group, d_group.data(), sizeof(int) * (size_t)nd) != 0)
      return -61;
    u64 *d_key = ptr<u64>(w64a_), *d_key_s = ptr<u64>(w64b_);
    u32 *d_p = ptr<u32>(w32a_), *d_p_s = ptr<u32>(w32b_);
    MakeKeysTiled mk{t_seq, nd, t_ws, t_begin, t_group, d_bin_start_, ptr<u32>(part_p_), ptr<u32>(slot_), table_.val, ptr<int>(a_list_), ptr<int>(b_list_),
                     ptr<int>(prod_a_), ptr<int>(prod_b_), ptr<int>(c_blkp_), lowbits, tilebits, tile_, (int)ntc, d_key, d_p};
    if (int rc = x_.for_each(nkept_, mk)) return rc;
    if (int rc = x_.sort_pairs(d_key, d_key_s, d_p, d_p_s, nkept_, lowbits + tilebits + gbits)) return rc;
    if (int rc = ensure(out3_, sizeof(int) * 3 * (size_t)nkept_)) return rc;
    WriteStack3 ws3{d_p_s, ptr<u32>(slot_), table_.val, ptr<int>(a_list_), ptr<int>(b_list_), ptr<int>(prod_a_), ptr<int>(prod_b_), ptr<int>(c_blkp_),
                    ptr<int>(out3_)};
    if (int rc = x_.for_each(nkept_, ws3)) return rc;
    // the new dispatch list
    std::vector<DevDispatch> nl;
    long long pos = 0;
    const int S = std::max(1, cfg.mm_stack_size);
    for (int sl = 0; sl < nslices; ++sl)
      for (int ws = 1; ws <= nstacks; ++ws) {
        long long left = gcount[(size_t)sl * (nstacks + 1) + ws];
        while (left > 0) {
          DevDispatch e;
          e.ws = ws;
          e.begin = -1;
          e.size = (int)std::min<long long>(left, S);
          e.slice = sl;
          e.seq_start = pos;
          nl.push_back(e);
          pos += e.size;
          left -= e.size;
        }
      }
    out.dispatch.swap(nl);
    tiled_last_ = true;
    return 0;
  }

  const int* stack3(const DevDispatch& d) const override { return ptr<int>(out3_) + 3 * (size_t)d.seq_start; }
  int fetch_stack3(const DevDispatch& d, int* host3) override {
    if (x_.d2h(host3, stack3(d), sizeof(int) * 3 * (size_t)d.size) != 0) return -61;
    return x_.sync();
  }
  int store_stack3(const DevDispatch& d, const int* host3) override {
    if (x_.h2d(const_cast<int*>(stack3(d)), host3, sizeof(int) * 3 * (size_t)d.size) != 0) return -61;
    return 0;  // pageable source: staged when the call returns
  }
  int fetch_params7(const DevDispatch& d, int* host7) override {
    if (int rc = ensure(p7_, sizeof(int) * 7 * (size_t)std::max(d.size, 1))) return rc;
    // host order of the stack: traversal order (reference mode) or the stack's own order (tile mode)
    const u32* plist = tiled_last_ ? ptr<u32>(w32b_) : ptr<u32>(part_p_);
    const long long offset = tiled_last_ ? d.seq_start : (long long)h_bin_start_[(size_t)d.ws] + d.begin;
    WriteParams7 f{maps_, plist, ptr<u32>(slot_), table_.val, ptr<int>(a_list_), ptr<int>(b_list_), ptr<int>(prod_a_),
                   ptr<int>(prod_b_), ptr<int>(c_blkp_), offset, ptr<int>(p7_)};
    if (int rc = x_.for_each(d.size, f)) return rc;
    if (x_.d2h(host7, p7_.p, sizeof(int) * 7 * (size_t)d.size) != 0) return -61;
    return x_.sync();
  }
  int fetch_index(int from, int n, int* rows, int* cols, int* blk_p) override {
    if (n <= 0) return 0;
    if (x_.d2h(rows, ptr<int>(c_row_) + from, sizeof(int) * (size_t)n) != 0 || x_.d2h(cols, ptr<int>(c_col_) + from, sizeof(int) * (size_t)n) != 0 ||
        x_.d2h(blk_p, ptr<int>(c_blkp_) + from, sizeof(int) * (size_t)n) != 0)
      return -61;
    return x_.sync();
  }
};

}  // namespace

IDeviceBuilder* make_device_builder(void* cuda_stream) { return new Builder<DeviceExec>(cuda_stream); }
IDeviceBuilder* make_emulated_builder() { return new Builder<HostExec>(nullptr); }

}  // namespace dbcsr_b200
