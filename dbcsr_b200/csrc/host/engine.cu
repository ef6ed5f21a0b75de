// dbcsr_b200/csrc/host/engine.cu -- host-side engine behind include/dbcsr_b200_host.h: per-thread multrec/csr stack builder
// feeding the accelerator through the acc/libsmm C ABI exactly the way DBCSR's accdrv does
// (src/mm/dbcsr_mm_accdrv.F:170-219 init, :433-541 process, :340-362 finalize): per-thread stream, ring of stack buffers
// (pinned + device) guarded by events, stack ordered by C, async H2D, libsmm_acc_process, device-resident C buffer.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <thread>
#include <numeric>
#include <vector>

#include "../../../include/dbcsr_acc.h"
#include "../../../include/dbcsr_acc_libsmm.h"
#include "../../../include/dbcsr_b200_host.h"
#include "stack_builder.hpp"
#include "device_builder.hpp"

using dbcsr_b200::Config;
using dbcsr_b200::Idx3;
using dbcsr_b200::LocalMultiply;
using dbcsr_b200::StackDescr;

namespace {

constexpr int kMaxKernelDim = 80;  // src/core/dbcsr_config.F:185

Config to_config(const dbcsr_b200_cfg_t& c) {
  Config k;
  k.mm_stack_size = c.mm_stack_size;
  k.n_stacks = c.n_stacks;
  k.multrec_limit = c.multrec_limit;
  k.stack_sort = c.stack_sort;
  k.min_flop_sort = c.min_flop_sort;
  k.binning_nbins = c.binning_nbins;
  k.binning_binsize = c.binning_binsize;
  return k;
}

struct RecordedStack {
  StackDescr d;
  int thread = 0, stack_number = 0, size = 0;
  std::vector<int> host, dev;
};

struct StackBuffer {  // stack_buffer_type, src/mm/dbcsr_mm_accdrv.F:63-72
  int* host = nullptr;   // pinned, 3 x mm_stack_size
  void* dev = nullptr;   // device, 3 x mm_stack_size
  void* calculated = nullptr;  // event
};

struct ThreadState {
  std::unique_ptr<LocalMultiply> mm;
  void* stream = nullptr;
  std::vector<StackBuffer> bufs;
  void* c_dev = nullptr;
  size_t c_capacity = 0;   // elements of the device C buffer
  size_t c_requested = 0;  // capacity asked for at creation (0 = estimate)
  size_t c_used_high = 0;  // largest datasize since the buffer was last zeroed: what reset has to clear
  // statistics of the host-driver route (stacks the accelerator refused): entries, stacks, flop
  long long cpu_entries = 0, cpu_stacks = 0, cpu_flop = 0;
  std::vector<RecordedStack> recorded;
  int a_first = 1, a_last = 0;
  std::vector<std::pair<int, int>> slices;  // (a_first, a_last) of every row chunk this thread owns, in row order
  double* c_host = nullptr;  // optional: D2H target enqueued right behind this thread's last stack
  bool has_preset = false;   // work matrix started from existing C blocks (beta != 0 / retain_sparsity)
  // result of dbcsr_b200_engine_filter_c: compacted index + compacted device data area
  bool filtered = false;
  std::vector<int> f_rows, f_cols, f_blk_p;
  int f_datasize = 0;
  void* c_final = nullptr;
  size_t c_final_capacity = 0;
  int rc = 0;
  double build_seconds = 0.0;
  // device-side builder (DBCSR_B200_DEVICE_BUILD): own stream for the index passes, joined into `stream` by an event
  std::unique_ptr<dbcsr_b200::IDeviceBuilder> devb;
  cudaStream_t build_stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // early D2H of finished row slices beside the stack kernels
  cudaEvent_t build_done = nullptr, drain_done = nullptr, slice_done = nullptr;
  bool dev_index = false;  // the product index of the current multiply lives in the device builder
  long long dev_ticks = 0;  // ticks built by the device passes (introspection)
  int dev_ticks_mult = 0;   // ... of the current multiply
  std::vector<int> h7, h3, idx_r, idx_c, idx_p;
  // statistics of dbcsr_mm_sched (src/mm/dbcsr_mm_sched.F:266-382,392-461): per (m,n,k) entries / stacks handed to the
  // accelerator and how many of those stacks ran on an untuned kernel; inhomogeneous stacks are booked under (0,0,0)
  struct MnkStat {
    int m = 0, n = 0, k = 0;
    long long entries_acc = 0, nstacks_acc = 0, nstacks_acc_untuned = 0, flop = 0;
  };
  std::vector<MnkStat> stats;
  void stats_add(int m, int n, int k, long long entries, long long flop, bool untuned) {
    for (auto& x : stats)
      if (x.m == m && x.n == n && x.k == k) {
        x.entries_acc += entries;
        x.nstacks_acc += 1;
        x.nstacks_acc_untuned += untuned ? 1 : 0;
        x.flop += flop;
        return;
      }
    MnkStat x;
    x.m = m, x.n = n, x.k = k, x.entries_acc = entries, x.nstacks_acc = 1, x.nstacks_acc_untuned = untuned ? 1 : 0, x.flop = flop;
    stats.push_back(x);
  }
};

}  // namespace

struct dbcsr_b200_engine {
  dbcsr_b200_cfg_t cfg;
  Config kcfg;
  int mode = 0;
  int device = 0;
  int nrows = 0, ncols = 0, nk = 0;
  std::vector<int> m_sizes, n_sizes, k_sizes;
  std::vector<ThreadState> th;
  std::vector<Idx3> a_sorted, b_sorted;
  // on-the-fly filter: per-row thresholds, norms permuted like the sorted lists
  std::vector<float> row_eps, a_norm_sorted, b_norm_sorted;
  std::vector<int> blk_tmp;
  // optional: events[c] = "the A rows of chunk c are on the device" (pipelined panel upload); consumed by the next multiply
  std::vector<void*> chunk_events;
  // host driver of the scheduler (dbcsr_mm_sched_process: a stack the accelerator refuses is processed by the CPU driver,
  // src/mm/dbcsr_mm_sched.F:340-363).  Supplied by the caller; this library contains no CPU compute path.
  dbcsr_b200_host_driver_fn host_driver = nullptr;
  void* host_driver_ctx = nullptr;
  size_t nnz_hint = 0;  // elements of the panels of the current multiply (initial size estimate of the C buffers)
};

extern "C" {

void dbcsr_b200_cfg_default(dbcsr_b200_cfg_t* cfg) {
  cfg->mm_stack_size = 30000;
  cfg->n_stacks = 3;
  cfg->multrec_limit = 512;
  cfg->stack_sort = 1;
  cfg->min_flop_sort = 4000;
  cfg->binning_nbins = 4096;
  cfg->binning_binsize = 16;
  cfg->thread_buffers = 8;
  cfg->row_chunks = 1;
  cfg->dev_tile = 0;
}

void dbcsr_b200_rec_sort_index(int nrows, int ncols, int nblks, int* list3) {
  static_assert(sizeof(Idx3) == 3 * sizeof(int), "Idx3 must be three packed ints");
  std::vector<Idx3> tmp((size_t)std::max(nblks, 0));
  if (nblks > 0) dbcsr_b200::rec_sort_index(1, nrows, 1, ncols, reinterpret_cast<Idx3*>(list3), nblks, tmp);
}
void dbcsr_b200_rec_sort_index_mt(int nrows, int ncols, int nblks, int* list3, int depth) {
  if (nblks > 0) dbcsr_b200::rec_sort_index_mt(1, nrows, 1, ncols, reinterpret_cast<Idx3*>(list3), nblks, depth);
}
void dbcsr_b200_stack_sort(const int* params7, int* out3, int stack_size) { dbcsr_b200::stack_sort(params7, out3, stack_size); }
void dbcsr_b200_stack_binning(const int* params7, int* out3, int stack_size, int nbins, int binsize) {
  dbcsr_b200::stack_binning(params7, out3, stack_size, nbins, binsize);
}

dbcsr_b200_engine_t* dbcsr_b200_engine_create(const dbcsr_b200_cfg_t* cfg, const int* m_sizes, int nrows, const int* n_sizes, int ncols,
                                              const int* k_sizes, int nk, int nthreads, int mode, size_t c_capacity) {
  if (cfg == nullptr || nthreads < 1 || nrows < 0 || ncols < 0 || nk < 0) return nullptr;
  auto* e = new dbcsr_b200_engine();
  e->cfg = *cfg;
  e->kcfg = to_config(*cfg);
  e->mode = mode;
  if (mode & DBCSR_B200_LAUNCH) cudaGetDevice(&e->device);
  e->nrows = nrows;
  e->ncols = ncols;
  e->nk = nk;
  e->m_sizes.assign(m_sizes, m_sizes + nrows);
  e->n_sizes.assign(n_sizes, n_sizes + ncols);
  e->k_sizes.assign(k_sizes, k_sizes + nk);
  e->th.resize((size_t)nthreads);
  for (int t = 0; t < nthreads; ++t) {
    ThreadState& ts = e->th[t];
    ts.mm.reset(new LocalMultiply(e->kcfg, e->m_sizes, e->n_sizes, e->k_sizes));
    if (mode & DBCSR_B200_LAUNCH) {
      // dbcsr_mm_accdrv_init (src/mm/dbcsr_mm_accdrv.F:170-219,279-307): stream, stack buffers; the C buffer is created
      // lazily at the first multiply, when the rows owned by the thread are known
      if (c_dbcsr_acc_stream_create(&ts.stream, "dbcsr_b200 thread", 0) != 0) goto fail;
      ts.bufs.resize((size_t)std::max(1, cfg->thread_buffers));
      for (auto& b : ts.bufs) {
        const size_t bytes = sizeof(int) * 3 * (size_t)cfg->mm_stack_size;
        if (c_dbcsr_acc_host_mem_allocate(reinterpret_cast<void**>(&b.host), bytes, ts.stream) != 0) goto fail;
        if (c_dbcsr_acc_dev_mem_allocate(&b.dev, bytes) != 0) goto fail;
        if (c_dbcsr_acc_event_create(&b.calculated) != 0) goto fail;
      }
      ts.c_requested = c_capacity;
    }
    if (mode & DBCSR_B200_DEVICE_BUILD) {
      if (mode & DBCSR_B200_LAUNCH) {
        if (cudaStreamCreateWithFlags(&ts.build_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ts.build_done, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ts.drain_done, cudaEventDisableTiming) != cudaSuccess ||
            cudaStreamCreateWithFlags(&ts.copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ts.slice_done, cudaEventDisableTiming) != cudaSuccess)
          goto fail;
        ts.devb.reset(dbcsr_b200::make_device_builder(ts.build_stream));
      }
      else {
        ts.devb.reset(dbcsr_b200::make_emulated_builder());
      }
      ts.devb->set_tile_order(cfg->dev_tile);
      ts.dev_index = true;
    }
  }
  return e;
fail:
  dbcsr_b200_engine_destroy(e);
  return nullptr;
}

void dbcsr_b200_engine_destroy(dbcsr_b200_engine_t* e) {
  if (e == nullptr) return;
  for (auto& ts : e->th) {
    if (ts.stream != nullptr) c_dbcsr_acc_stream_sync(ts.stream);
    for (auto& b : ts.bufs) {
      if (b.host != nullptr) c_dbcsr_acc_host_mem_deallocate(b.host, ts.stream);
      if (b.dev != nullptr) c_dbcsr_acc_dev_mem_deallocate(b.dev);
      if (b.calculated != nullptr) c_dbcsr_acc_event_destroy(b.calculated);
    }
    if (ts.c_dev != nullptr) c_dbcsr_acc_dev_mem_deallocate(ts.c_dev);
    if (ts.c_final != nullptr) c_dbcsr_acc_dev_mem_deallocate(ts.c_final);
    ts.devb.reset();
    if (ts.build_done != nullptr) cudaEventDestroy(ts.build_done);
    if (ts.drain_done != nullptr) cudaEventDestroy(ts.drain_done);
    if (ts.slice_done != nullptr) cudaEventDestroy(ts.slice_done);
    if (ts.copy_stream != nullptr) {
      cudaStreamSynchronize(ts.copy_stream);
      cudaStreamDestroy(ts.copy_stream);
    }
    if (ts.build_stream != nullptr) cudaStreamDestroy(ts.build_stream);
    if (ts.stream != nullptr) c_dbcsr_acc_stream_destroy(ts.stream);
  }
  delete e;
}

// device C buffers (dbcsr_mm_accdrv_init, src/mm/dbcsr_mm_accdrv.F:170-219: sized like the work area, zeroed asynchronously);
// created at first use.  Initial size: the requested capacity, else the dense upper bound of the thread's block rows when all
// threads' bounds together fit into a quarter of the free device memory (then the buffer never grows), else an estimate from
// the panels' element counts; grow_c_buffer enlarges it on demand like dbcsr_data_ensure_size (src/mm/dbcsr_mm_accdrv.F:471-473).
constexpr size_t kMaxOffsets = 0x7fffffffull;  // offsets are int32 (hard part 8)

static size_t dense_bound(const dbcsr_b200_engine_t* e, int t) {
  const int nthreads = (int)e->th.size();
  size_t sum_n = 0;
  for (int v : e->n_sizes) sum_n += (size_t)v;
  const int rcn = std::max(1, e->cfg.row_chunks);
  const int nchunks = nthreads * rcn;
  size_t sum_m = 0;
  for (int c = t; c < nchunks; c += nthreads) {  // ALL block rows this thread owns (later Cannon ticks may touch rows that have
                                                 // no A block in the first panel)
    const int row_lo = (int)(((long long)e->nrows * c) / nchunks), row_hi = (int)(((long long)e->nrows * (c + 1)) / nchunks);
    for (int r = row_lo; r < row_hi; ++r) sum_m += (size_t)e->m_sizes[(size_t)r];
  }
  return sum_m * sum_n;
}

static int ensure_c_buffers(dbcsr_b200_engine_t* e) {
  if (!(e->mode & DBCSR_B200_LAUNCH)) return 0;
  const int nthreads = (int)e->th.size();
  size_t dense_total = 0;
  for (int t = 0; t < nthreads; ++t) dense_total += std::min(dense_bound(e, t), kMaxOffsets);
  size_t free_b = 0, total_b = 0;
  const bool dense_fits = c_dbcsr_acc_dev_mem_info(&free_b, &total_b) == 0 && dense_total * sizeof(double) <= free_b / 4;
  for (int t = 0; t < nthreads; ++t) {
    ThreadState& ts = e->th[t];
    if (ts.c_dev != nullptr) continue;
    size_t cap = ts.c_requested;
    if (cap == 0) {
      const size_t dense = dense_bound(e, t);
      // sparse products: start from twice the operands' share of this thread and grow on demand
      cap = dense_fits ? dense : std::min(dense, std::max<size_t>(2 * e->nnz_hint / (size_t)nthreads, (size_t)1 << 20));
    }
    if (cap == 0) cap = 1;
    if (cap > kMaxOffsets) cap = kMaxOffsets;
    if (c_dbcsr_acc_dev_mem_allocate(&ts.c_dev, cap * sizeof(double)) != 0) return -40;
    ts.c_capacity = cap;
    ts.c_used_high = 0;
    if (c_dbcsr_acc_memset_zero(ts.c_dev, 0, cap * sizeof(double), ts.stream) != 0) return -41;
  }
  return 0;
}

// dbcsr_data_ensure_size(c_buffer, datasize, factor) of the accelerator driver: a larger buffer, the old contents copied device
// to device behind everything enqueued on the thread's stream, the new tail zeroed; the old buffer is released after the stream
// has drained (stack kernels and downloads that still use it are all on this stream).
static int grow_c_buffer(ThreadState& ts, size_t need) {
  if (need > kMaxOffsets) return -42;
  size_t cap = std::max(need + need / 2, ts.c_capacity * 2);
  if (cap > kMaxOffsets) cap = kMaxOffsets;
  void* bigger = nullptr;
  if (c_dbcsr_acc_dev_mem_allocate(&bigger, cap * sizeof(double)) != 0) {
    cap = need;  // second try: exactly what is needed
    if (c_dbcsr_acc_dev_mem_allocate(&bigger, cap * sizeof(double)) != 0) return -40;
  }
  if (c_dbcsr_acc_memcpy_d2d(ts.c_dev, bigger, ts.c_capacity * sizeof(double), ts.stream) != 0) return -41;
  if (c_dbcsr_acc_memset_zero(bigger, ts.c_capacity * sizeof(double), (cap - ts.c_capacity) * sizeof(double), ts.stream) != 0) return -41;
  if (c_dbcsr_acc_stream_sync(ts.stream) != 0) return -41;
  if (c_dbcsr_acc_dev_mem_deallocate(ts.c_dev) != 0) return -41;
  ts.c_dev = bigger;
  ts.c_capacity = cap;
  return 0;
}

// owner thread of a (1-based) block row: chunk c = rows (c*nrows/C, (c+1)*nrows/C], C = nthreads*row_chunks, thread = c mod nthreads
static int owner_of_row(const dbcsr_b200_engine_t* e, int row) {
  const int nthreads = (int)e->th.size();
  const int nchunks = nthreads * std::max(1, e->cfg.row_chunks);
  int c = (int)(((long long)(row - 1) * nchunks) / std::max(1, e->nrows));
  while (c > 0 && row <= (int)(((long long)e->nrows * c) / nchunks)) --c;
  while (c < nchunks - 1 && row > (int)(((long long)e->nrows * (c + 1)) / nchunks)) ++c;
  return c % nthreads;
}

}  // extern "C"

// One tick of host thread t with the device-side builder (device_builder.hpp): all row slices of the thread in one pass, then
// the stacks in the host builder's dispatch order.  Stacks the accelerator driver sorts by c_first arrive in device order; the
// others (binning, inhomogeneous) are ordered on the host from their 7-wide entries, exactly like the host path.
template <class BookFn>
static int device_build_tick(dbcsr_b200_engine_t* e, int t, const void* a_dev, const void* b_dev, int nb, bool filter, const BookFn& book) {
  ThreadState& ts = e->th[(size_t)t];
  const bool launch = (e->mode & DBCSR_B200_LAUNCH) != 0;
  const int na = (int)e->a_sorted.size();
  if (launch && ts.drain_done != nullptr) {
    // the previous tick's stack kernels read the stack array this build overwrites
    if (cudaEventRecord(ts.drain_done, *static_cast<cudaStream_t*>(ts.stream)) != cudaSuccess ||
        cudaStreamWaitEvent(ts.build_stream, ts.drain_done, 0) != cudaSuccess)
      return -45;
  }
  dbcsr_b200::DevBuildResult r;
  dbcsr_b200::DevBuildOptions opt;
  if (filter) {  // on-the-fly filter: norms follow the sorted lists (src/mm/dbcsr_mm_csr.F:270-278)
    opt.a_norms = e->a_norm_sorted.data();
    opt.b_norms = e->b_norm_sorted.data();
    opt.row_eps = e->row_eps.data();
  }
  opt.c_sym = ts.mm->c_symmetry();
  opt.global_rows = ts.mm->c_global_rows().empty() ? nullptr : ts.mm->c_global_rows().data();
  opt.global_cols = ts.mm->c_global_cols().empty() ? nullptr : ts.mm->c_global_cols().data();
  opt.keep_sparsity = ts.mm->keep_sparsity();
  const int rc_b = ts.devb->build(*ts.mm, e->a_sorted.data(), na, ts.slices, e->b_sorted.data(), nb, r, opt);
  if (rc_b != 0) return rc_b;
  // new part of the C index -> host work index (the engine's accessors, finalize and the downloads use it)
  const int nnew = r.nblk_after - r.nblk_before;
  ts.idx_r.resize((size_t)std::max(nnew, 1));
  ts.idx_c.resize((size_t)std::max(nnew, 1));
  ts.idx_p.resize((size_t)std::max(nnew, 1));
  // (LAUNCH: fetched behind the stack launches -- the kernels do not need the host copy of the index)
  if (!launch)
    if (int rc = ts.devb->fetch_index(r.nblk_before, nnew, ts.idx_r.data(), ts.idx_c.data(), ts.idx_p.data())) return rc;
  long long flop = 0;
  const size_t S7 = 7 * (size_t)std::max(1, e->kcfg.mm_stack_size);
  // stacks ordered on the host
  for (const auto& d : r.dispatch) {
    const StackDescr& sd = ts.mm->descr(d.ws);
    const bool on_dev = dbcsr_b200::IDeviceBuilder::device_ordered(e->kcfg, sd);
    if (on_dev && !(e->mode & DBCSR_B200_RECORD)) continue;
    ts.h7.resize(std::max(ts.h7.size(), S7));
    ts.h3.resize(std::max(ts.h3.size(), S7));
    if (int rc = ts.devb->fetch_params7(d, ts.h7.data())) return rc;
    if (!on_dev) {
      dbcsr_b200::accdrv_order_stack(e->kcfg, sd, ts.h7.data(), ts.h3.data(), d.size, true);
      if (int rc = ts.devb->store_stack3(d, ts.h3.data())) return rc;
    }
    if (e->mode & DBCSR_B200_RECORD) {
      RecordedStack rec;
      rec.d = sd;
      rec.thread = t;
      rec.stack_number = d.ws;
      rec.size = d.size;
      rec.host.assign(ts.h7.begin(), ts.h7.begin() + 7 * (size_t)d.size);
      rec.dev.resize(3 * (size_t)d.size);
      if (on_dev) {
        if (int rc = ts.devb->fetch_stack3(d, rec.dev.data())) return rc;
      }
      else {
        std::copy(ts.h3.begin(), ts.h3.begin() + 3 * (size_t)d.size, rec.dev.begin());
      }
      ts.recorded.push_back(std::move(rec));
    }
  }
  for (const auto& d : r.dispatch) {
    const StackDescr& sd = ts.mm->descr(d.ws);
    if (sd.defined_mnk) flop += 2LL * sd.m * sd.n * sd.k * d.size;
  }
  // inhomogeneous stacks: flop from their entries (only these need the host copy again)
  if (!launch) {
    ts.mm->append_index(ts.idx_r.data(), ts.idx_c.data(), ts.idx_p.data(), nnew, r.datasize_after, 0);
    for (const auto& d : r.dispatch) {
      const StackDescr& sd = ts.mm->descr(d.ws);
      if (!sd.defined_mnk) {
        ts.h7.resize(std::max(ts.h7.size(), S7));
        if (int rc = ts.devb->fetch_params7(d, ts.h7.data())) return rc;
        for (int i = 0; i < d.size; ++i) flop += 2LL * ts.h7[7 * (size_t)i] * ts.h7[7 * (size_t)i + 1] * ts.h7[7 * (size_t)i + 2];
      }
      book(sd, ts.h7.data(), d.size, false);
    }
    ts.mm->append_index(nullptr, nullptr, nullptr, 0, r.datasize_after, flop);
    return 0;
  }
  // ---- LAUNCH: C buffer large enough, stack kernels behind the build
  if ((size_t)r.datasize_after > ts.c_capacity) {
    if (ts.copy_stream != nullptr) cudaStreamSynchronize(ts.copy_stream);  // downloads of earlier slices read the old buffer
    if (int grc = grow_c_buffer(ts, (size_t)r.datasize_after)) return grc;
  }
  ts.c_used_high = std::max(ts.c_used_high, (size_t)r.datasize_after);
  cudaStream_t st = *static_cast<cudaStream_t*>(ts.stream);
  if (cudaEventRecord(ts.build_done, ts.build_stream) != cudaSuccess || cudaStreamWaitEvent(st, ts.build_done, 0) != cudaSuccess) return -45;
  const int nslices = (int)ts.slices.size();
  size_t di = 0;
  int ds_prev = r.datasize_before;
  for (int s = 0; s < nslices; ++s) {
    const int chunk = t + s * (int)e->th.size();
    if (chunk < (int)e->chunk_events.size() && e->chunk_events[(size_t)chunk] != nullptr) {
      if (c_dbcsr_acc_stream_wait_event(ts.stream, e->chunk_events[(size_t)chunk]) != 0) return -49;
    }
    for (; di < r.dispatch.size() && r.dispatch[di].slice == s; ++di) {
      const auto& d = r.dispatch[di];
      const StackDescr& sd = ts.mm->descr(d.ws);
      const int* host7 = nullptr;
      if (!sd.defined_mnk) {  // the library bins an inhomogeneous stack from its host entries
        ts.h7.resize(std::max(ts.h7.size(), S7));
        if (int rc = ts.devb->fetch_params7(d, ts.h7.data())) return rc;
        host7 = ts.h7.data();
        for (int i = 0; i < d.size; ++i) flop += 2LL * host7[7 * (size_t)i] * host7[7 * (size_t)i + 1] * host7[7 * (size_t)i + 2];
      }
      const int rc = libsmm_acc_process(host7, ts.devb->stack3(d), d.size, dbcsr_type_real_8, a_dev, b_dev, ts.c_dev, sd.max_m, sd.max_n, sd.max_k,
                                        kMaxKernelDim, sd.defined_mnk, ts.stream, ts.stream);
      if (rc < 0) {
        if (e->host_driver == nullptr) return rc;
        if (host7 == nullptr) {
          ts.h7.resize(std::max(ts.h7.size(), S7));
          if (int frc = ts.devb->fetch_params7(d, ts.h7.data())) return frc;
          host7 = ts.h7.data();
        }
        const int hrc = e->host_driver(e->host_driver_ctx, t, sd.m, sd.n, sd.k, sd.defined_mnk, host7, d.size, r.datasize_after);
        if (hrc != 0) return hrc < 0 ? hrc : -50;
        long long f = 0;
        for (int i = 0; i < d.size; ++i) f += 2LL * host7[7 * (size_t)i] * host7[7 * (size_t)i + 1] * host7[7 * (size_t)i + 2];
        ts.cpu_entries += d.size;
        ts.cpu_stacks += 1;
        ts.cpu_flop += f;
        continue;
      }
      book(sd, host7, d.size, rc == 10);
    }
    const int ds1 = r.slice_datasize[(size_t)s];
    if (ts.c_host != nullptr && ts.c_dev != nullptr && ds1 > ds_prev && !ts.has_preset) {
      // the C blocks created by this slice are final once its stacks have drained: download them on the copy stream while the
      // stacks of the next slice run
      if (cudaEventRecord(ts.slice_done, st) != cudaSuccess || cudaStreamWaitEvent(ts.copy_stream, ts.slice_done, 0) != cudaSuccess ||
          cudaMemcpyAsync(ts.c_host + ds_prev, static_cast<double*>(ts.c_dev) + ds_prev, (size_t)(ds1 - ds_prev) * sizeof(double),
                          cudaMemcpyDeviceToHost, ts.copy_stream) != cudaSuccess)
        return -46;
    }
    ds_prev = ds1;
  }
  if (ts.has_preset && ts.c_host != nullptr && ts.c_dev != nullptr && r.datasize_after > 0) {
    // products accumulate into the pre-existing blocks too: one D2H of the whole work area behind the last stack
    if (c_dbcsr_acc_memcpy_d2h(ts.c_dev, ts.c_host, (size_t)r.datasize_after * sizeof(double), ts.stream) != 0) return -46;
  }
  // everything is enqueued: now the host copy of the new part of the C index (build stream; the stacks run meanwhile)
  if (int rc = ts.devb->fetch_index(r.nblk_before, nnew, ts.idx_r.data(), ts.idx_c.data(), ts.idx_p.data())) return rc;
  ts.mm->append_index(ts.idx_r.data(), ts.idx_c.data(), ts.idx_p.data(), nnew, r.datasize_after, flop);
  return 0;
}

extern "C" {

static int engine_multiply_impl(dbcsr_b200_engine_t* e, const int* a_list3, int na, const void* a_dev, const float* a_norms,
                                const int* b_list3, int nb, const void* b_dev, const float* b_norms) {
  if (e == nullptr || na < 0 || nb < 0) return -1;
  const int nthreads = (int)e->th.size();
  const bool filter = a_norms != nullptr && b_norms != nullptr && !e->row_eps.empty();
  for (auto& ts : e->th) ts.filtered = false;
  {
    size_t nnz = 0;
    for (int i = 0; i < na; ++i)
      nnz += (size_t)e->m_sizes[(size_t)a_list3[3 * (size_t)i] - 1] * (size_t)e->th[0].mm->k_size(a_list3[3 * (size_t)i + 1]);
    for (int i = 0; i < nb; ++i)
      nnz += (size_t)e->th[0].mm->k_size(b_list3[3 * (size_t)i]) * (size_t)e->n_sizes[(size_t)b_list3[3 * (size_t)i + 1] - 1];
    e->nnz_hint = nnz;
  }
  // --- left panel: split the BCSR-ordered list over the threads by block rows (DBCSR: thr_c slices of coo_l, each slice
  //     rec-sorted on its own with the full panel extents, src/mm/dbcsr_mm_cannon.F:2910-2967)
  e->a_sorted.resize((size_t)na);
  std::memcpy(e->a_sorted.data(), a_list3, sizeof(int) * 3 * (size_t)na);
  e->b_sorted.resize((size_t)nb);
  std::memcpy(e->b_sorted.data(), b_list3, sizeof(int) * 3 * (size_t)nb);
  if (filter) {
    // the norms must follow their blocks through rec_sort_index: sort with the list position in the blk slot, then put the
    // element offsets back (the reference computes the norms after the sort, src/mm/dbcsr_mm_cannon.F:1127-1160)
    for (int i = 0; i < na; ++i) e->a_sorted[(size_t)i].blk = i;
    for (int i = 0; i < nb; ++i) e->b_sorted[(size_t)i].blk = i;
  }
  {
    // static ownership: chunk c = block rows (c*nrows/C, (c+1)*nrows/C], C = nthreads * row_chunks; thread t owns the chunks
    // t, t+T, t+2T, ... in every tick, so that the C rows of different threads stay disjoint over a whole Cannon multiply;
    // the BCSR-ordered list is sorted by row => contiguous slices.  row_chunks = 1 is DBCSR's one slice per thread.
    const int rc = std::max(1, e->cfg.row_chunks);
    const int nchunks = nthreads * rc;
    for (auto& ts : e->th) ts.slices.clear();
    int pos = 0;
    for (int c = 0; c < nchunks; ++c) {
      const int row_hi = (int)(((long long)e->nrows * (c + 1)) / nchunks);
      int end = pos;
      while (end < na && e->a_sorted[(size_t)end].row <= row_hi) ++end;
      if (c == nchunks - 1) end = na;
      e->th[(size_t)(c % nthreads)].slices.emplace_back(pos + 1, end);
      pos = end;
    }
    for (auto& ts : e->th) {  // kept for the dense-bound computation below: overall span of the thread's slices
      ts.a_first = ts.slices.front().first;
      ts.a_last = ts.slices.back().second;
    }
  }
  // --- sort panels
  {
    std::vector<std::thread> workers;
    for (int t = 0; t < nthreads; ++t) {
      workers.emplace_back([e, t]() {
        ThreadState& ts = e->th[t];
        std::vector<Idx3> tmp;
        for (const auto& sl : ts.slices) {
          const int cnt = sl.second - sl.first + 1;
          if (cnt > 0) dbcsr_b200::rec_sort_index(1, e->nrows, 1, e->nk, e->a_sorted.data() + (sl.first - 1), cnt, tmp);
        }
      });
    }
    if (nb > 0) {
      // the right panel is ONE list: its sort would be the serial part of every multiply (10 ms for the 1e5 blocks of config 2)
      int depth = 0;
      // with the device-side builder the engine runs few host threads (one per build pipeline): the sort still uses the cores
      const int sort_threads = (e->mode & DBCSR_B200_DEVICE_BUILD) ? std::max(nthreads, (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()))) : nthreads;
      while ((1 << depth) < sort_threads && depth < 4) ++depth;
      dbcsr_b200::rec_sort_index_mt(1, e->nk, 1, e->ncols, e->b_sorted.data(), nb, depth);
    }
    for (auto& w : workers) w.join();
  }
  if (filter) {
    e->a_norm_sorted.resize((size_t)na);
    e->b_norm_sorted.resize((size_t)nb);
    for (int j = 0; j < na; ++j) {
      const int i = e->a_sorted[(size_t)j].blk;
      e->a_norm_sorted[(size_t)j] = a_norms[i];
      e->a_sorted[(size_t)j].blk = a_list3[3 * (size_t)i + 2];
    }
    for (int j = 0; j < nb; ++j) {
      const int i = e->b_sorted[(size_t)j].blk;
      e->b_norm_sorted[(size_t)j] = b_norms[i];
      e->b_sorted[(size_t)j].blk = b_list3[3 * (size_t)i + 2];
    }
  }
  {
    const int rc_alloc = ensure_c_buffers(e);
    if (rc_alloc != 0) return rc_alloc;
  }
  // --- per-thread multrec -> csr -> sched -> accdrv
  std::vector<std::thread> workers;
  for (int t = 0; t < nthreads; ++t) {
    workers.emplace_back([e, t, a_dev, b_dev, nb, filter]() {
      ThreadState& ts = e->th[t];
      ts.rc = 0;
      if (e->mode & DBCSR_B200_LAUNCH) cudaSetDevice(e->device);  // the active device is per host thread
      const auto t0 = std::chrono::steady_clock::now();
      int next_buf = 0;
      auto book = [&](const StackDescr& d, const int* params7, int size, bool untuned) {
        long long flop = 0;
        if (d.defined_mnk) {
          flop = 2LL * d.m * d.n * d.k * size;
        }
        else {
          for (int i = 0; i < size; ++i) flop += 2LL * params7[7 * (size_t)i] * params7[7 * (size_t)i + 1] * params7[7 * (size_t)i + 2];
        }
        ts.stats_add(d.defined_mnk ? d.m : 0, d.defined_mnk ? d.n : 0, d.defined_mnk ? d.k : 0, size, flop, untuned);
      };
      auto dispatch = [&](int stack_number, const StackDescr& d, const int* params7, int size) {
        if (ts.rc != 0) return;
        if (!(e->mode & DBCSR_B200_LAUNCH)) book(d, params7, size, false);
        if (e->mode & DBCSR_B200_RECORD) {
          RecordedStack r;
          r.d = d;
          r.thread = t;
          r.stack_number = stack_number;
          r.size = size;
          r.host.assign(params7, params7 + 7 * (size_t)size);
          r.dev.resize(3 * (size_t)size);
          dbcsr_b200::accdrv_order_stack(e->kcfg, d, params7, r.dev.data(), size, true);
          ts.recorded.push_back(std::move(r));
        }
        if (!(e->mode & DBCSR_B200_LAUNCH)) return;
        if ((size_t)ts.mm->datasize() > ts.c_capacity) {  // new C blocks beyond the buffer: grow it (src/mm/dbcsr_mm_accdrv.F:471-473)
          const int grc = grow_c_buffer(ts, (size_t)ts.mm->datasize());
          if (grc != 0) {
            ts.rc = grc;
            return;
          }
        }
        ts.c_used_high = std::max(ts.c_used_high, (size_t)ts.mm->datasize());
        // pick a stack buffer whose previous kernel has finished (round robin + event wait instead of the reference's busy poll)
        StackBuffer& b = ts.bufs[(size_t)next_buf];
        next_buf = (next_buf + 1) % (int)ts.bufs.size();
        if (c_dbcsr_acc_event_synchronize(b.calculated) != 0) {
          ts.rc = -43;
          return;
        }
        dbcsr_b200::accdrv_order_stack(e->kcfg, d, params7, b.host, size, true);
        if (c_dbcsr_acc_memcpy_h2d(b.host, b.dev, sizeof(int) * 3 * (size_t)size, ts.stream) != 0) {
          ts.rc = -44;
          return;
        }
        const int rc = libsmm_acc_process(params7, static_cast<const int*>(b.dev), size, dbcsr_type_real_8, a_dev, b_dev, ts.c_dev,
                                          d.max_m, d.max_n, d.max_k, kMaxKernelDim, d.defined_mnk, ts.stream, ts.stream);
        if (rc < 0) {
          // The accelerator refused the stack and left C untouched.  dbcsr_mm_sched_process now hands the stack to the host driver
          // (src/mm/dbcsr_mm_sched.F:340-363), whose contributions live in the HOST work matrix and are added to the downloaded
          // device buffer at finalize (src/mm/dbcsr_mm_accdrv.F:340-362).  This library has no CPU compute path: the route exists
          // only when the caller installed a host driver (dbcsr_b200_engine_set_host_driver); otherwise the multiply fails loudly.
          if (e->host_driver == nullptr) {
            ts.rc = rc;
            return;
          }
          const int hrc = e->host_driver(e->host_driver_ctx, t, d.m, d.n, d.k, d.defined_mnk, params7, size, ts.mm->datasize());
          if (hrc != 0) {
            ts.rc = hrc < 0 ? hrc : -50;
            return;
          }
          long long flop = 0;
          for (int i = 0; i < size; ++i) flop += 2LL * params7[7 * (size_t)i] * params7[7 * (size_t)i + 1] * params7[7 * (size_t)i + 2];
          ts.cpu_entries += size;
          ts.cpu_stacks += 1;
          ts.cpu_flop += flop;
          return;
        }
        book(d, params7, size, rc == 10);
        if (c_dbcsr_acc_event_record(b.calculated, ts.stream) != 0) ts.rc = -45;
      };
      // ---- device-side builder (a multiply uses one builder from reset to reset: dev_index is true after create / reset and
      //      cleared when the device passes cannot handle the engine's configuration, which shows at the first tick)
      if (ts.devb != nullptr && ts.dev_index) {
        const int rc_dev = device_build_tick(e, t, a_dev, b_dev, nb, filter, book);
        if (rc_dev == 0) {
          ++ts.dev_ticks;
          ++ts.dev_ticks_mult;
        }
        // -60: more stacks than the device passes handle; -40: no device memory for the passes' work arrays.  Before the first
        // device-built tick of a multiply nothing has been committed, so the host builder below can take the whole multiply
        if (!(rc_dev == -60 || (rc_dev == -40 && ts.dev_ticks_mult == 0))) {
          if (rc_dev != 0 && ts.rc == 0) ts.rc = rc_dev;
          ts.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
          return;
        }
        ts.dev_index = false;
      }
      int slice_no = 0;
      for (const auto& sl : ts.slices) {
        const int chunk = t + (slice_no++) * (int)e->th.size();
        if ((e->mode & DBCSR_B200_LAUNCH) && chunk < (int)e->chunk_events.size() && e->chunk_events[(size_t)chunk] != nullptr && ts.rc == 0) {
          // the stacks of this row chunk read A blocks that may still be in flight: order them behind the chunk's upload
          if (c_dbcsr_acc_stream_wait_event(ts.stream, e->chunk_events[(size_t)chunk]) != 0) ts.rc = -49;
        }
        const size_t ds0 = (size_t)ts.mm->datasize();
        ts.mm->multiply(e->a_sorted.data(), sl.first, sl.second, e->b_sorted.data(), nb, dispatch,  // ends with a purge
                        filter ? e->a_norm_sorted.data() : nullptr, filter ? e->b_norm_sorted.data() : nullptr);
        const size_t ds1 = (size_t)ts.mm->datasize();
        if (ts.rc == 0 && ts.c_host != nullptr && ts.c_dev != nullptr && ds1 > ds0 && !ts.has_preset) {
          // the C blocks created by this row chunk are final once the stream drains: start their D2H now, while this thread
          // builds its next chunk and the other threads are still busy
          if (c_dbcsr_acc_memcpy_d2h(static_cast<double*>(ts.c_dev) + ds0, ts.c_host + ds0, (ds1 - ds0) * sizeof(double), ts.stream) != 0)
            ts.rc = -46;
        }
      }
      if (ts.rc == 0 && ts.has_preset && ts.c_host != nullptr && ts.c_dev != nullptr && ts.mm->datasize() > 0) {
        // products accumulate into the pre-existing blocks too: one D2H of the whole work area behind the last stack
        if (c_dbcsr_acc_memcpy_d2h(ts.c_dev, ts.c_host, (size_t)ts.mm->datasize() * sizeof(double), ts.stream) != 0) ts.rc = -46;
      }
      ts.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    });
  }
  for (auto& w : workers) w.join();
  e->chunk_events.clear();  // one-shot
  for (auto& ts : e->th)
    if (ts.rc != 0) return ts.rc;
  return 0;
}

int dbcsr_b200_engine_multiply(dbcsr_b200_engine_t* e, const int* a_list3, int na, const void* a_dev, const int* b_list3, int nb,
                               const void* b_dev) {
  return engine_multiply_impl(e, a_list3, na, a_dev, nullptr, b_list3, nb, b_dev, nullptr);
}

int dbcsr_b200_engine_multiply_filtered(dbcsr_b200_engine_t* e, const int* a_list3, int na, const void* a_dev, const float* a_norms,
                                        const int* b_list3, int nb, const void* b_dev, const float* b_norms) {
  if (a_norms == nullptr || b_norms == nullptr) return -1;
  return engine_multiply_impl(e, a_list3, na, a_dev, a_norms, b_list3, nb, b_dev, b_norms);
}

int dbcsr_b200_engine_set_filter(dbcsr_b200_engine_t* e, const float* row_max_epss) {
  if (e == nullptr) return -1;
  if (row_max_epss == nullptr)
    e->row_eps.clear();
  else
    e->row_eps.assign(row_max_epss, row_max_epss + e->nrows);
  for (auto& ts : e->th) ts.mm->set_row_eps(e->row_eps);
  return 0;
}

void dbcsr_b200_row_max_epss(double filter_eps, const int* total_row_counts, int nrows, float* row_max_epss) {
  // src/mm/dbcsr_mm_cannon.F:1098-1107: everything in single precision
  const float eps_sp = (float)filter_eps;
  for (int r = 0; r < nrows; ++r) {
    const float q = eps_sp / (float)std::max(1, total_row_counts[r]);
    row_max_epss[r] = q * q;
  }
}

int dbcsr_b200_engine_preset_c(dbcsr_b200_engine_t* e, const int* rows, const int* cols, int nblks, const double* host_data,
                               int keep_sparsity) {
  // The work matrices start from the existing C blocks (dbcsr_mm_csr_init -> fill_hash_tables, src/mm/dbcsr_mm_csr.F:526-576):
  // block i goes to the thread that owns its block row, in list order, at the next free offset of that thread's work area.
  if (e == nullptr || nblks < 0) return -1;
  const int nthreads = (int)e->th.size();
  for (auto& ts : e->th)
    if (!ts.mm->c_row().empty()) return -2;  // must come before the first tick of a multiply (after create / reset)
  std::vector<std::vector<int>> r((size_t)nthreads), c((size_t)nthreads), p((size_t)nthreads);
  std::vector<std::vector<double>> data((size_t)nthreads);
  std::vector<long long> ds((size_t)nthreads, 0);
  size_t src = 0;
  for (int i = 0; i < nblks; ++i) {
    if (rows[i] < 1 || rows[i] > e->nrows || cols[i] < 1 || cols[i] > e->ncols) return -3;
    const int t = owner_of_row(e, rows[i]);
    const size_t nze = (size_t)e->m_sizes[(size_t)rows[i] - 1] * (size_t)e->n_sizes[(size_t)cols[i] - 1];
    r[(size_t)t].push_back(rows[i]);
    c[(size_t)t].push_back(cols[i]);
    p[(size_t)t].push_back((int)(ds[(size_t)t] + 1));
    ds[(size_t)t] += (long long)nze;
    if (ds[(size_t)t] > 0x7fffffffLL) return -4;
    if (host_data != nullptr && (e->mode & DBCSR_B200_LAUNCH)) data[(size_t)t].insert(data[(size_t)t].end(), host_data + src, host_data + src + nze);
    src += nze;
  }
  const int rc_alloc = ensure_c_buffers(e);
  if (rc_alloc != 0) return rc_alloc;
  for (int t = 0; t < nthreads; ++t) {
    ThreadState& ts = e->th[(size_t)t];
    ts.mm->preset_c(r[(size_t)t].data(), c[(size_t)t].data(), p[(size_t)t].data(), (int)r[(size_t)t].size(), (int)ds[(size_t)t]);
    ts.mm->set_keep_sparsity(keep_sparsity != 0);
    ts.has_preset = true;
    if (ts.devb != nullptr && ts.dev_index) {
      const int prc = ts.devb->preset(e->nrows, e->ncols, r[(size_t)t].data(), c[(size_t)t].data(), p[(size_t)t].data(), (int)r[(size_t)t].size(),
                                      (int)ds[(size_t)t]);
      if (prc != 0) return prc;
    }
    if ((e->mode & DBCSR_B200_LAUNCH) && !data[(size_t)t].empty()) {
      if ((size_t)ds[(size_t)t] > ts.c_capacity) {
        const int grc = grow_c_buffer(ts, (size_t)ds[(size_t)t]);
        if (grc != 0) return grc;
      }
      ts.c_used_high = std::max(ts.c_used_high, (size_t)ds[(size_t)t]);
      if (c_dbcsr_acc_memcpy_h2d(data[(size_t)t].data(), ts.c_dev, data[(size_t)t].size() * sizeof(double), ts.stream) != 0) return -44;
    }
  }
  for (int t = 0; t < nthreads; ++t)  // the staging vectors die here
    if (e->th[(size_t)t].stream != nullptr && c_dbcsr_acc_stream_sync(e->th[(size_t)t].stream) != 0) return -1;
  return 0;
}

int dbcsr_b200_engine_stats(const dbcsr_b200_engine_t* e, long long* table, int max_rows, long long* totals) {
  // the table DBCSR prints at finalize (dbcsr_mm_sched_print_statistics): merged over the threads like
  // stats_collect_from_threads (src/mm/dbcsr_mm_sched.F:463-505); accumulates over all multiplies of the engine
  if (e == nullptr) return -1;
  std::vector<ThreadState::MnkStat> all;
  for (const auto& ts : e->th)
    for (const auto& x : ts.stats) {
      bool found = false;
      for (auto& y : all)
        if (y.m == x.m && y.n == x.n && y.k == x.k) {
          y.entries_acc += x.entries_acc;
          y.nstacks_acc += x.nstacks_acc;
          y.nstacks_acc_untuned += x.nstacks_acc_untuned;
          y.flop += x.flop;
          found = true;
          break;
        }
      if (!found) all.push_back(x);
    }
  std::sort(all.begin(), all.end(), [](const ThreadState::MnkStat& a, const ThreadState::MnkStat& b) { return a.flop > b.flop; });
  long long tot_flop = 0, tot_entries = 0, tot_stacks = 0;
  for (const auto& x : all) {
    tot_flop += x.flop;
    tot_entries += x.entries_acc;
    tot_stacks += x.nstacks_acc;
  }
  if (totals != nullptr) {
    totals[0] = tot_flop;
    totals[1] = tot_entries;
    totals[2] = tot_stacks;
  }
  const int n = (int)all.size();
  if (table != nullptr)
    for (int i = 0; i < n && i < max_rows; ++i) {
      long long* r = table + 7 * (size_t)i;
      r[0] = all[(size_t)i].m;
      r[1] = all[(size_t)i].n;
      r[2] = all[(size_t)i].k;
      r[3] = all[(size_t)i].entries_acc;
      r[4] = all[(size_t)i].nstacks_acc;
      r[5] = all[(size_t)i].nstacks_acc_untuned;
      r[6] = all[(size_t)i].flop;
    }
  return n;
}

int dbcsr_b200_engine_set_host_driver(dbcsr_b200_engine_t* e, dbcsr_b200_host_driver_fn fn, void* ctx) {
  if (e == nullptr) return -1;
  e->host_driver = fn;
  e->host_driver_ctx = ctx;
  return 0;
}

int dbcsr_b200_engine_stats_cpu(const dbcsr_b200_engine_t* e, long long* totals) {
  if (e == nullptr || totals == nullptr) return -1;
  totals[0] = totals[1] = totals[2] = 0;
  for (const auto& ts : e->th) {
    totals[0] += ts.cpu_flop;
    totals[1] += ts.cpu_entries;
    totals[2] += ts.cpu_stacks;
  }
  return 0;
}

int dbcsr_b200_engine_set_c_symmetry(dbcsr_b200_engine_t* e, int on, const int* global_rows, const int* global_cols) {
  // product with symmetry: skip the half of the off-diagonal blocks that the checkerboard rule stores transposed
  // (src/mm/dbcsr_mm_csr.F:280-292; c_local_rows / c_local_cols of the reference = global_rows / global_cols here)
  if (e == nullptr) return -1;
  std::vector<int> gr, gc;
  if (global_rows != nullptr) gr.assign(global_rows, global_rows + e->nrows);
  if (global_cols != nullptr) gc.assign(global_cols, global_cols + e->ncols);
  for (auto& ts : e->th) ts.mm->set_c_symmetry(on != 0, gr, gc);
  return 0;
}

int dbcsr_b200_filter_index(double filter_eps, const double* norms2, int nblks, int* rows, int* cols, int* blk_p, const int* nelems,
                            long long* nze_after) {
  // multrec_filtering_d, src/mm/dbcsr_mm_multrec.F:700-758: keep a block iff DDOT(blk, blk) >= filter_eps**2; kept entries move
  // to the front in order, their blk_p values are unchanged (the data area keeps its holes)
  const double eps2 = filter_eps * filter_eps;
  int last = 0;
  long long nze = 0;
  for (int b = 0; b < nblks; ++b) {
    if (blk_p[b] == 0 || nelems[b] == 0) continue;
    if (norms2[b] >= eps2) {
      if (last < b) {
        rows[last] = rows[b];
        cols[last] = cols[b];
        blk_p[last] = blk_p[b];
      }
      ++last;
      nze += nelems[b];
    }
  }
  if (nze_after != nullptr) *nze_after = nze;
  return last;
}

int dbcsr_b200_finalize_index(int nblks, int* rows, int* cols, const int* nelems, int* perm, int* blk_p_new, long long* nze) {
  // dbcsr_finalize / dbcsr_merge_all, index part (work/dbcsr_work_operations.F:749-958): the work index (order of first touch)
  // becomes BCSR order -- rows ascending, columns ascending within a row -- and the data area is compacted in that order.
  if (nblks < 0) return -1;
  std::vector<int> order((size_t)nblks);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
    if (rows[x] != rows[y]) return rows[x] < rows[y];
    return cols[x] < cols[y];
  });
  std::vector<int> r((size_t)nblks), c((size_t)nblks);
  long long run = 0;
  for (int i = 0; i < nblks; ++i) {
    const int o = order[(size_t)i];
    r[(size_t)i] = rows[o];
    c[(size_t)i] = cols[o];
    perm[i] = o;
    if (run + 1 > 0x7fffffffLL) return -4;
    blk_p_new[i] = (int)(run + 1);
    run += nelems[o];
  }
  for (int i = 1; i < nblks; ++i)
    if (r[(size_t)i] == r[(size_t)i - 1] && c[(size_t)i] == c[(size_t)i - 1]) return -5;  // a block may exist once only
  std::copy(r.begin(), r.end(), rows);
  std::copy(c.begin(), c.end(), cols);
  if (nze != nullptr) *nze = run;
  return 0;
}

namespace {
// Final filter (do_filter; dbcsr_mm_multrec_finalize -> multrec_filtering) and/or BCSR ordering (bcsr_order; dbcsr_finalize) of
// every thread's product BEFORE the download: squared block norms in double on the device, 8 B per block to the host, index
// compaction / sort there, then the kept blocks are gathered into a contiguous device area in index order -- only those travel
// over PCIe, and they arrive in their final place.
int finalize_impl(dbcsr_b200_engine_t* e, double filter_eps, bool do_filter, bool bcsr_order) {
  if (e == nullptr || !(e->mode & DBCSR_B200_LAUNCH)) return -1;
  struct Scratch {
    std::vector<int> offs, nel;
    std::vector<double> norms;
    void *d_offs = nullptr, *d_nel = nullptr, *d_norms = nullptr, *d_dst = nullptr;
  };
  const int nthreads = (int)e->th.size();
  std::vector<Scratch> sc((size_t)nthreads);
  int rc = 0;
  auto release = [&]() {
    for (auto& x : sc)
      for (void* d : {x.d_offs, x.d_nel, x.d_norms, x.d_dst})
        if (d != nullptr) c_dbcsr_acc_dev_mem_deallocate(d);
  };
  // phase 1: norms of every thread's blocks, all streams busy at once
  for (int t = 0; t < nthreads && rc == 0; ++t) {
    ThreadState& ts = e->th[(size_t)t];
    Scratch& x = sc[(size_t)t];
    ts.filtered = false;
    const int nb = (int)ts.mm->c_row().size();
    if (nb == 0 || ts.c_dev == nullptr) continue;
    x.offs.resize((size_t)nb);
    x.nel.resize((size_t)nb);
    x.norms.resize((size_t)nb);
    for (int b = 0; b < nb; ++b) {
      x.offs[(size_t)b] = ts.mm->c_blk_p()[(size_t)b] - 1;
      x.nel[(size_t)b] = e->m_sizes[(size_t)ts.mm->c_row()[(size_t)b] - 1] * e->n_sizes[(size_t)ts.mm->c_col()[(size_t)b] - 1];
    }
    const size_t ib = sizeof(int) * (size_t)nb;
    if (c_dbcsr_acc_dev_mem_allocate(&x.d_offs, ib) != 0 || c_dbcsr_acc_dev_mem_allocate(&x.d_nel, ib) != 0 ||
        c_dbcsr_acc_dev_mem_allocate(&x.d_dst, ib) != 0 || c_dbcsr_acc_dev_mem_allocate(&x.d_norms, sizeof(double) * (size_t)nb) != 0) {
      rc = -40;
      break;
    }
    if (do_filter &&
        (c_dbcsr_acc_memcpy_h2d(x.offs.data(), x.d_offs, ib, ts.stream) != 0 || c_dbcsr_acc_memcpy_h2d(x.nel.data(), x.d_nel, ib, ts.stream) != 0 ||
         libsmm_acc_b200_block_norms_f64(static_cast<const double*>(ts.c_dev), nb, static_cast<const int*>(x.d_offs),
                                         static_cast<const int*>(x.d_nel), static_cast<double*>(x.d_norms), ts.stream) != 0 ||
         c_dbcsr_acc_memcpy_d2h(x.d_norms, x.norms.data(), sizeof(double) * (size_t)nb, ts.stream) != 0))
      rc = -47;
  }
  // phase 2: compaction of the index on the host, gather of the kept blocks on the device
  for (int t = 0; t < nthreads && rc == 0; ++t) {
    ThreadState& ts = e->th[(size_t)t];
    Scratch& x = sc[(size_t)t];
    const int nb = (int)ts.mm->c_row().size();
    ts.f_rows = ts.mm->c_row();
    ts.f_cols = ts.mm->c_col();
    ts.f_blk_p = ts.mm->c_blk_p();
    ts.f_datasize = 0;
    if (nb == 0 || ts.c_dev == nullptr) {
      ts.filtered = true;
      continue;
    }
    if (c_dbcsr_acc_stream_sync(ts.stream) != 0) {
      rc = -1;
      break;
    }
    long long nze = 0;
    int kept = nb;
    if (do_filter)
      kept = dbcsr_b200_filter_index(filter_eps, x.norms.data(), nb, ts.f_rows.data(), ts.f_cols.data(), ts.f_blk_p.data(), x.nel.data(),
                                     &nze);
    ts.f_rows.resize((size_t)kept);
    ts.f_cols.resize((size_t)kept);
    ts.f_blk_p.resize((size_t)kept);
    if (bcsr_order && kept > 1) {  // rows ascending, columns ascending within a row (dbcsr_finalize); blk_p follows its block
      std::vector<int> order((size_t)kept);
      std::iota(order.begin(), order.end(), 0);
      std::stable_sort(order.begin(), order.end(), [&](int p, int q) {
        if (ts.f_rows[(size_t)p] != ts.f_rows[(size_t)q]) return ts.f_rows[(size_t)p] < ts.f_rows[(size_t)q];
        return ts.f_cols[(size_t)p] < ts.f_cols[(size_t)q];
      });
      std::vector<int> r2((size_t)kept), c2((size_t)kept), p2((size_t)kept);
      for (int j = 0; j < kept; ++j) {
        r2[(size_t)j] = ts.f_rows[(size_t)order[(size_t)j]];
        c2[(size_t)j] = ts.f_cols[(size_t)order[(size_t)j]];
        p2[(size_t)j] = ts.f_blk_p[(size_t)order[(size_t)j]];
      }
      ts.f_rows.swap(r2);
      ts.f_cols.swap(c2);
      ts.f_blk_p.swap(p2);
    }
    // gather plan: kept block j moves from its work-area offset to the next free offset of the compact area
    std::vector<int>& src_off = x.offs;  // reuse
    std::vector<int> dst_off((size_t)kept);
    int run = 0;
    for (int j = 0; j < kept; ++j) {
      src_off[(size_t)j] = ts.f_blk_p[(size_t)j] - 1;
      x.nel[(size_t)j] = e->m_sizes[(size_t)ts.f_rows[(size_t)j] - 1] * e->n_sizes[(size_t)ts.f_cols[(size_t)j] - 1];
      dst_off[(size_t)j] = run;
      ts.f_blk_p[(size_t)j] = run + 1;
      run += x.nel[(size_t)j];
    }
    ts.f_datasize = run;
    if ((size_t)run > ts.c_final_capacity) {
      if (ts.c_final != nullptr) c_dbcsr_acc_dev_mem_deallocate(ts.c_final);
      ts.c_final = nullptr;
      ts.c_final_capacity = 0;
      if (c_dbcsr_acc_dev_mem_allocate(&ts.c_final, (size_t)run * sizeof(double)) != 0) {
        rc = -40;
        break;
      }
      ts.c_final_capacity = (size_t)run;
    }
    if (kept > 0) {
      const size_t kb = sizeof(int) * (size_t)kept;
      if (c_dbcsr_acc_memcpy_h2d(src_off.data(), x.d_offs, kb, ts.stream) != 0 || c_dbcsr_acc_memcpy_h2d(x.nel.data(), x.d_nel, kb, ts.stream) != 0 ||
          c_dbcsr_acc_memcpy_h2d(dst_off.data(), x.d_dst, kb, ts.stream) != 0 ||
          libsmm_acc_b200_gather_blocks(static_cast<const double*>(ts.c_dev), static_cast<double*>(ts.c_final), kept,
                                        static_cast<const int*>(x.d_offs), static_cast<const int*>(x.d_dst), static_cast<const int*>(x.d_nel),
                                        ts.stream) != 0) {
        rc = -48;
        break;
      }
      // dst_off lives in this scope only: the H2D from pageable memory has been staged when memcpy_h2d returns, but make the
      // lifetime explicit
      if (c_dbcsr_acc_stream_sync(ts.stream) != 0) {
        rc = -1;
        break;
      }
    }
    ts.filtered = true;
  }
  for (auto& ts : e->th)
    if (ts.stream != nullptr) c_dbcsr_acc_stream_sync(ts.stream);
  release();
  return rc;
}
}  // namespace

int dbcsr_b200_engine_filter_c(dbcsr_b200_engine_t* e, double filter_eps) { return finalize_impl(e, filter_eps, true, false); }

int dbcsr_b200_engine_finalize_c(dbcsr_b200_engine_t* e, double filter_eps) {
  // dbcsr_finalize on the device: optional final filter (filter_eps >= 0), then every thread's blocks in BCSR order, compacted
  return finalize_impl(e, filter_eps, filter_eps >= 0.0, true);
}

int dbcsr_b200_engine_nchunks(const dbcsr_b200_engine_t* e) {
  return e == nullptr ? 0 : (int)e->th.size() * std::max(1, e->cfg.row_chunks);
}

int dbcsr_b200_engine_chunk_rows(const dbcsr_b200_engine_t* e, int chunk, int* row_lo, int* row_hi) {
  const int nchunks = dbcsr_b200_engine_nchunks(e);
  if (e == nullptr || chunk < 0 || chunk >= nchunks) return -1;
  *row_lo = (int)(((long long)e->nrows * chunk) / nchunks);        // chunk = block rows (row_lo, row_hi], 1-based
  *row_hi = (int)(((long long)e->nrows * (chunk + 1)) / nchunks);
  return 0;
}

int dbcsr_b200_engine_set_chunk_events(dbcsr_b200_engine_t* e, void* const* events, int nevents) {
  if (e == nullptr || nevents < 0 || nevents > dbcsr_b200_engine_nchunks(e)) return -1;
  e->chunk_events.assign(events, events + nevents);
  return 0;
}

int dbcsr_b200_engine_set_k_sizes(dbcsr_b200_engine_t* e, const int* k_sizes, int nk) {
  if (e == nullptr || nk < 0) return -1;
  std::vector<int> ks(k_sizes, k_sizes + nk);
  e->nk = nk;
  for (auto& ts : e->th) ts.mm->set_k_sizes(ks);
  return 0;
}

int dbcsr_b200_engine_reset(dbcsr_b200_engine_t* e) {
  // start a new multiply on pooled resources (streams, stack buffers, device C buffer stay allocated like DBCSR's memory pools,
  // src/data/dbcsr_mem_methods.F:41-251): forget the product index, clear recorded stacks, zero the C buffer asynchronously
  if (e == nullptr) return -1;
  for (auto& ts : e->th) {
    ts.recorded.clear();
    ts.has_preset = false;
    ts.filtered = false;
    // only what the last multiply touched needs clearing (everything behind it is still zero)
    const size_t used = std::min(std::max(ts.c_used_high, (size_t)ts.mm->datasize()), ts.c_capacity);
    ts.mm->reset();
    ts.mm->set_k_sizes(e->k_sizes);
    if (ts.copy_stream != nullptr && cudaStreamSynchronize(ts.copy_stream) != cudaSuccess) return -41;
    if (ts.devb != nullptr && ts.devb->reset() != 0) return -41;
    ts.dev_index = ts.devb != nullptr;
    ts.dev_ticks_mult = 0;
    if (ts.c_dev != nullptr && used > 0 && c_dbcsr_acc_memset_zero(ts.c_dev, 0, used * sizeof(double), ts.stream) != 0) return -41;
    ts.c_used_high = 0;
  }
  return 0;
}

int dbcsr_b200_engine_sync(dbcsr_b200_engine_t* e) {
  if (e == nullptr) return -1;
  for (auto& ts : e->th) {
    if (ts.stream != nullptr && c_dbcsr_acc_stream_sync(ts.stream) != 0) return -1;
    if (ts.copy_stream != nullptr && cudaStreamSynchronize(ts.copy_stream) != cudaSuccess) return -1;
  }
  return 0;
}

long long dbcsr_b200_engine_device_built_ticks(const dbcsr_b200_engine_t* e) {
  long long n = 0;
  if (e != nullptr)
    for (const auto& ts : e->th) n += ts.dev_ticks;
  return n;
}

int dbcsr_b200_engine_nthreads(const dbcsr_b200_engine_t* e) { return (int)e->th.size(); }
int dbcsr_b200_engine_c_nblks(const dbcsr_b200_engine_t* e, int t) {
  const ThreadState& ts = e->th[(size_t)t];
  return ts.filtered ? (int)ts.f_rows.size() : (int)ts.mm->c_row().size();
}
int dbcsr_b200_engine_c_datasize(const dbcsr_b200_engine_t* e, int t) {
  const ThreadState& ts = e->th[(size_t)t];
  return ts.filtered ? ts.f_datasize : ts.mm->datasize();
}
const int* dbcsr_b200_engine_c_rows(const dbcsr_b200_engine_t* e, int t) {
  const ThreadState& ts = e->th[(size_t)t];
  return ts.filtered ? ts.f_rows.data() : ts.mm->c_row().data();
}
const int* dbcsr_b200_engine_c_cols(const dbcsr_b200_engine_t* e, int t) {
  const ThreadState& ts = e->th[(size_t)t];
  return ts.filtered ? ts.f_cols.data() : ts.mm->c_col().data();
}
const int* dbcsr_b200_engine_c_blk_p(const dbcsr_b200_engine_t* e, int t) {
  const ThreadState& ts = e->th[(size_t)t];
  return ts.filtered ? ts.f_blk_p.data() : ts.mm->c_blk_p().data();
}
void* dbcsr_b200_engine_c_dev(const dbcsr_b200_engine_t* e, int t) {
  const ThreadState& ts = e->th[(size_t)t];
  return ts.filtered ? ts.c_final : ts.c_dev;
}

int dbcsr_b200_engine_set_c_host(dbcsr_b200_engine_t* e, int t, double* host) {
  if (e == nullptr || t < 0 || t >= (int)e->th.size()) return -1;
  e->th[(size_t)t].c_host = host;
  return 0;
}

size_t dbcsr_b200_engine_c_capacity(const dbcsr_b200_engine_t* e, int t) { return e->th[(size_t)t].c_capacity; }

int dbcsr_b200_engine_wait_event(dbcsr_b200_engine_t* e, void* event) {
  // every thread stream waits for `event` (e.g. "panels uploaded and transposed") before running anything enqueued later
  if (e == nullptr || event == nullptr) return -1;
  for (auto& ts : e->th)
    if (ts.stream != nullptr && c_dbcsr_acc_stream_wait_event(ts.stream, event) != 0) return -1;
  return 0;
}

int dbcsr_b200_engine_c_to_host_async(dbcsr_b200_engine_t* e, int t, double* host) {
  // D2H of the thread's C buffer enqueued behind its last stack; completes at dbcsr_b200_engine_sync
  ThreadState& ts = e->th[(size_t)t];
  if (ts.c_dev == nullptr || ts.stream == nullptr) return -1;
  const size_t n = ts.filtered ? (size_t)ts.f_datasize : (size_t)ts.mm->datasize();
  const void* src = ts.filtered ? ts.c_final : ts.c_dev;
  if (n > 0 && c_dbcsr_acc_memcpy_d2h(src, host, n * sizeof(double), ts.stream) != 0) return -1;
  return 0;
}

int dbcsr_b200_engine_c_to_host(dbcsr_b200_engine_t* e, int t, double* host) {
  ThreadState& ts = e->th[(size_t)t];
  if (ts.c_dev == nullptr || ts.stream == nullptr) return -1;
  const size_t n = ts.filtered ? (size_t)ts.f_datasize : (size_t)ts.mm->datasize();
  const void* src = ts.filtered ? ts.c_final : ts.c_dev;
  if (n > 0 && c_dbcsr_acc_memcpy_d2h(src, host, n * sizeof(double), ts.stream) != 0) return -1;
  return c_dbcsr_acc_stream_sync(ts.stream);
}

long long dbcsr_b200_engine_flop(const dbcsr_b200_engine_t* e) {
  long long f = 0;
  for (auto& ts : e->th) f += ts.mm->flop();
  return f;
}

double dbcsr_b200_engine_build_seconds(const dbcsr_b200_engine_t* e) {
  double s = 0;
  for (auto& ts : e->th) s = std::max(s, ts.build_seconds);
  return s;
}

int dbcsr_b200_engine_nstacks(const dbcsr_b200_engine_t* e) {
  size_t n = 0;
  for (auto& ts : e->th) n += ts.recorded.size();
  return (int)n;
}

static const RecordedStack* find_stack(const dbcsr_b200_engine_t* e, int i) {
  for (auto& ts : e->th) {
    if (i < (int)ts.recorded.size()) return &ts.recorded[(size_t)i];
    i -= (int)ts.recorded.size();
  }
  return nullptr;
}

void dbcsr_b200_engine_stack_info(const dbcsr_b200_engine_t* e, int i, int* info) {
  const RecordedStack* r = find_stack(e, i);
  if (r == nullptr) return;
  info[0] = r->d.m;
  info[1] = r->d.n;
  info[2] = r->d.k;
  info[3] = r->d.max_m;
  info[4] = r->d.max_n;
  info[5] = r->d.max_k;
  info[6] = r->d.defined_mnk;
  info[7] = r->size;
  info[8] = r->thread;
  info[9] = r->stack_number;
}
const int* dbcsr_b200_engine_stack_host(const dbcsr_b200_engine_t* e, int i) {
  const RecordedStack* r = find_stack(e, i);
  return r ? r->host.data() : nullptr;
}
const int* dbcsr_b200_engine_stack_dev(const dbcsr_b200_engine_t* e, int i) {
  const RecordedStack* r = find_stack(e, i);
  return r ? r->dev.data() : nullptr;
}

static int transpose_panel_impl(const int* b_list3, int nb, const int* k_sizes, const int* n_sizes, void* b_dev, int* scratch_host,
                                void* scratch_dev, float* dev_norms, void* stream) {
  // group the blocks by (k,n) size pair, 0-based offsets (blk_p - 1), one transpose stack per pair
  // (src/mm/dbcsr_mm_common.F:346-496); with dev_norms the second half of the scratch carries the list position of every block
  if (nb <= 0) return 0;
  std::vector<int> order((size_t)nb);
  for (int i = 0; i < nb; ++i) order[(size_t)i] = i;
  auto key = [&](int i) {
    const int k = k_sizes[b_list3[3 * (size_t)i] - 1], n = n_sizes[b_list3[3 * (size_t)i + 1] - 1];
    return ((long long)k << 32) | (unsigned)n;
  };
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return key(x) < key(y); });
  for (int i = 0; i < nb; ++i) scratch_host[i] = b_list3[3 * (size_t)order[(size_t)i] + 2] - 1;
  const bool fused = dev_norms != nullptr;
  if (fused)
    for (int i = 0; i < nb; ++i) scratch_host[(size_t)nb + i] = order[(size_t)i];
  if (c_dbcsr_acc_memcpy_h2d(scratch_host, scratch_dev, sizeof(int) * (size_t)nb * (fused ? 2 : 1), stream) != 0) return -1;
  const int* d_offs = static_cast<const int*>(scratch_dev);
  int start = 0;
  while (start < nb) {
    int end = start;
    const long long k0 = key(order[(size_t)start]);
    while (end < nb && key(order[(size_t)end]) == k0) ++end;
    const int k = (int)(k0 >> 32), n = (int)(k0 & 0xffffffff);
    int rc;
    if (fused && k <= kMaxKernelDim && n <= kMaxKernelDim) {
      rc = libsmm_acc_b200_transpose_norms(d_offs, d_offs + nb, start, end - start, static_cast<double*>(b_dev), k, n, kMaxKernelDim, dev_norms,
                                           stream);
    }
    else {
      rc = libsmm_acc_transpose(d_offs, start, end - start, b_dev, dbcsr_type_real_8, k, n, kMaxKernelDim, stream);
      if (rc == 0 && fused) rc = -3;  // blocks above max_kernel_dim are not transposed: the caller computes their norms separately
    }
    if (rc != 0) return rc;
    start = end;
  }
  return 0;
}

int dbcsr_b200_transpose_panel(const int* b_list3, int nb, const int* k_sizes, const int* n_sizes, void* b_dev, int* scratch_host,
                               void* scratch_dev, void* stream) {
  return transpose_panel_impl(b_list3, nb, k_sizes, n_sizes, b_dev, scratch_host, scratch_dev, nullptr, stream);
}

int dbcsr_b200_transpose_panel_norms(const int* b_list3, int nb, const int* k_sizes, const int* n_sizes, void* b_dev, int* scratch_host,
                                     void* scratch_dev, float* dev_norms, void* stream) {
  if (dev_norms == nullptr) return -2;
  return transpose_panel_impl(b_list3, nb, k_sizes, n_sizes, b_dev, scratch_host, scratch_dev, dev_norms, stream);
}

}  // extern "C"
