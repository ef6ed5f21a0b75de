// dbcsr_b200/csrc/host/record_engine.cpp -- the host stack builder WITHOUT any accelerator: sorts the panels, splits the left list
// over `nthreads` row slices like the engine (engine.cu; DBCSR: one OpenMP thread per slice, src/mm/dbcsr_mm_multrec.F:306-311), runs
// LocalMultiply per slice and keeps every dispatched 7-wide stack.  Pure C++ (no CUDA): built into its own small library,
// libdbcsr_b200_hostbuilder.so, so that the CPU reference arm of bench.py -- which times the oracle's restatement of
// blas_process_mm_stack_d on DBCSR-order stacks -- does not need the accelerator library at all.
#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include "stack_builder.hpp"

using dbcsr_b200::Config;
using dbcsr_b200::Idx3;
using dbcsr_b200::LocalMultiply;
using dbcsr_b200::StackDescr;

namespace {
struct Rec {
  StackDescr d;
  int thread = 0, stack_number = 0, size = 0;
  std::vector<int> host;
};
}  // namespace

struct dbcsr_b200_recorder {
  std::vector<std::vector<Rec>> per_thread;
  std::vector<int> datasize;
  std::vector<const Rec*> flat;
  long long flop = 0;
};

extern "C" {

dbcsr_b200_recorder* dbcsr_b200_recorder_run(const int* m_sizes, int nrows, const int* n_sizes, int ncols, const int* k_sizes, int nk,
                                             const int* a_list3, int na, const int* b_list3, int nb, int nthreads, int mm_stack_size,
                                             int n_stacks) {
  if (nthreads < 1 || nrows < 0 || ncols < 0 || nk < 0 || na < 0 || nb < 0) return nullptr;
  Config cfg;
  cfg.mm_stack_size = mm_stack_size;
  cfg.n_stacks = n_stacks;
  const std::vector<int> ms(m_sizes, m_sizes + nrows), ns(n_sizes, n_sizes + ncols), ks(k_sizes, k_sizes + nk);
  std::vector<Idx3> a((size_t)na), b((size_t)nb);
  if (na) std::memcpy(a.data(), a_list3, sizeof(int) * 3 * (size_t)na);
  if (nb) std::memcpy(b.data(), b_list3, sizeof(int) * 3 * (size_t)nb);
  // slices: thread t owns block rows (t*nrows/T, (t+1)*nrows/T]; the BCSR-ordered list is sorted by row => contiguous slices
  std::vector<std::pair<int, int>> slices((size_t)nthreads);
  int pos = 0;
  for (int t = 0; t < nthreads; ++t) {
    const int row_hi = (int)(((long long)nrows * (t + 1)) / nthreads);
    int end = pos;
    while (end < na && a[(size_t)end].row <= row_hi) ++end;
    if (t == nthreads - 1) end = na;
    slices[(size_t)t] = {pos + 1, end};
    pos = end;
  }
  if (nb > 0) {
    int depth = 0;
    while ((1 << depth) < nthreads && depth < 4) ++depth;
    dbcsr_b200::rec_sort_index_mt(1, nk, 1, ncols, b.data(), nb, depth);
  }
  auto* r = new dbcsr_b200_recorder();
  r->per_thread.resize((size_t)nthreads);
  r->datasize.assign((size_t)nthreads, 0);
  std::vector<long long> flops((size_t)nthreads, 0);
  std::vector<std::thread> workers;
  for (int t = 0; t < nthreads; ++t) {
    workers.emplace_back([&, t]() {
      const auto sl = slices[(size_t)t];
      const int cnt = sl.second - sl.first + 1;
      std::vector<Idx3> tmp;
      if (cnt > 0) dbcsr_b200::rec_sort_index(1, nrows, 1, nk, a.data() + (sl.first - 1), cnt, tmp);
      LocalMultiply mm(cfg, ms, ns, ks);
      auto& out = r->per_thread[(size_t)t];
      LocalMultiply::DispatchFn dispatch = [&](int stack_number, const StackDescr& d, const int* params7, int size) {
        Rec rec;
        rec.d = d;
        rec.thread = t;
        rec.stack_number = stack_number;
        rec.size = size;
        rec.host.assign(params7, params7 + 7 * (size_t)size);
        out.push_back(std::move(rec));
      };
      mm.multiply(a.data(), sl.first, sl.second, b.data(), nb, dispatch);
      r->datasize[(size_t)t] = mm.datasize();
      flops[(size_t)t] = mm.flop();
    });
  }
  for (auto& w : workers) w.join();
  for (int t = 0; t < nthreads; ++t) {
    r->flop += flops[(size_t)t];
    for (const Rec& rec : r->per_thread[(size_t)t]) r->flat.push_back(&rec);
  }
  return r;
}

int dbcsr_b200_recorder_nstacks(const dbcsr_b200_recorder* r) { return r == nullptr ? 0 : (int)r->flat.size(); }
// info: m, n, k, defined_mnk, size, thread, stack_number
void dbcsr_b200_recorder_stack_info(const dbcsr_b200_recorder* r, int i, int* info) {
  const Rec& x = *r->flat[(size_t)i];
  info[0] = x.d.m, info[1] = x.d.n, info[2] = x.d.k, info[3] = x.d.defined_mnk, info[4] = x.size, info[5] = x.thread, info[6] = x.stack_number;
}
const int* dbcsr_b200_recorder_stack_host(const dbcsr_b200_recorder* r, int i) { return r->flat[(size_t)i]->host.data(); }
int dbcsr_b200_recorder_datasize(const dbcsr_b200_recorder* r, int t) { return r->datasize[(size_t)t]; }
long long dbcsr_b200_recorder_flop(const dbcsr_b200_recorder* r) { return r->flop; }
void dbcsr_b200_recorder_free(dbcsr_b200_recorder* r) { delete r; }

}  // extern "C"
