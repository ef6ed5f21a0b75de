/*
 * tools/kbench.c -- kernel-variant bench of the FP64 stack drain on ONE GPU, pure C on the drop-in ABI (no Python start-up, so a
 * whole sweep fits into a short GPU slot).  Workload = BASELINE.json configs[1] in structure: nblk x nblk block grid, 23x23 blocks,
 * Bernoulli(occ) block presence (own LCG, seeded), stacks built by the library's host engine (multrec order, MM_STACK_SIZE 30000,
 * C-sorted) exactly as bench.py does, operands resident on the device.
 *
 *   kbench <library.so> <out_dir> <nblk> <occ> <steps> <bsz | m,n,k> <spec> [<spec> ...]
 *   spec = variant:balance:chunk[:t]   (balance: bit 0 = balanced chunks, bit 1 = run-aligned chunk boundaries, negative = the
 *                                       library's per-shape policy; chunk: entries per warp, 0 = one wave, negative = policy; 0:-1:-1 is
 *                                       what ships; t = record a kernel timeline of launches 100..102 of one step into out_dir)
 *
 * Per spec: (1) parity -- C is zeroed, every stack is drained once, the per-block sums of squares (libsmm_acc_b200_block_norms_f64)
 * are compared EXACTLY with those of the first spec (operands are small integers, so every summation order gives the same
 * doubles); (2) timing -- `steps` repetitions of the whole drain (334 launches), wall clock around stream synchronisation.
 * KBENCH_ACC_LIB=<other.so>: take the acc/libsmm ABI functions (init, streams, memory, libsmm_acc_process) from that library instead,
 * e.g. baseline/_ref/libdbcsr_acc_ref.so = the reference's own CUDA backend built for this GPU (baseline/Makefile): the same
 * stacks, the same parity check (block norms by the main library's kernel) and the same timing loop give the same-box GPU baseline.
 * This is a development tool: bench.py is the bench contract.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/dbcsr_b200_host.h"

typedef int (*fn_i_v)(void);
typedef int (*fn_set_dev)(int);
typedef int (*fn_stream_create)(void**, const char*, int);
typedef int (*fn_stream_sync)(void*);
typedef int (*fn_dev_alloc)(void**, size_t);
typedef int (*fn_dev_free)(void*);
typedef int (*fn_memcpy)(const void*, void*, size_t, void*);
typedef int (*fn_memset)(void*, size_t, size_t, void*);
typedef int (*fn_process)(const int*, const int*, int, int, const void*, const void*, void*, int, int, int, int, int, void*, void*);
typedef int (*fn_norms)(const double*, int, const int*, const int*, double*, void*);
typedef int (*fn_set_tun)(const char*, long long);
typedef long long (*fn_get_tun)(const char*);
typedef void (*fn_set_trace)(void*);
typedef void (*fn_cfg_default)(dbcsr_b200_cfg_t*);
typedef dbcsr_b200_engine_t* (*fn_eng_create)(const dbcsr_b200_cfg_t*, const int*, int, const int*, int, const int*, int, int, int, size_t);
typedef int (*fn_eng_multiply)(dbcsr_b200_engine_t*, const int*, int, const void*, const int*, int, const void*);
typedef int (*fn_eng_int)(const dbcsr_b200_engine_t*);
typedef int (*fn_eng_int_t)(const dbcsr_b200_engine_t*, int);
typedef const int* (*fn_eng_ptr_t)(const dbcsr_b200_engine_t*, int);
typedef void (*fn_eng_info)(const dbcsr_b200_engine_t*, int, int*);
typedef long long (*fn_eng_flop)(const dbcsr_b200_engine_t*);
typedef void (*fn_eng_destroy)(dbcsr_b200_engine_t*);

static void* lib;
static void* acc_lib; /* library providing the acc/libsmm ABI (== lib unless KBENCH_ACC_LIB is set) */
static void* sym_from(void* l, const char* name) {
  void* p = dlsym(l, name);
  if (p == NULL) {
    fprintf(stderr, "kbench: missing symbol %s\n", name);
    exit(2);
  }
  return p;
}
static int noop_set_tun(const char* n, long long v) { (void)n; (void)v; return 0; }
static void* sym(const char* name) {
  void* p = dlsym(lib, name);
  if (p == NULL) {
    fprintf(stderr, "kbench: missing symbol %s\n", name);
    exit(2);
  }
  return p;
}
static double now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
static uint64_t lcg_state = 88172645463325252ull;
static inline double lcg_uniform(void) {  /* xorshift64* */
  lcg_state ^= lcg_state >> 12;
  lcg_state ^= lcg_state << 25;
  lcg_state ^= lcg_state >> 27;
  return (double)((lcg_state * 2685821657736338717ull) >> 11) * (1.0 / 9007199254740992.0);
}
/* operand values: integers 0..3, a function of the element index (so the host check needs no copy of the panels) */
#define A_VAL(i) ((double)((((uint64_t)(i)) * 2654435761ull >> 7) & 3))
#define B_VAL(i) ((double)(((((uint64_t)(i)) + 12345) * 40503ull >> 5) & 3))
#define CHECK(x)                                                      \
  do {                                                                \
    int rc_ = (x);                                                    \
    if (rc_ != 0) {                                                   \
      fprintf(stderr, "kbench: %s failed with %d (line %d)\n", #x, rc_, __LINE__); \
      exit(3);                                                        \
    }                                                                 \
  } while (0)

/* BCSR-ordered list index (row, col, blk_p) of a random nblk x nblk pattern; returns the number of blocks */
static int make_pattern(int nblk, double occ, int blk_elems, int** list3_out) {
  int cap = (int)(nblk * (double)nblk * occ * 1.2) + 1024, n = 0;
  int* l = (int*)malloc(sizeof(int) * 3 * (size_t)cap);
  for (int r = 1; r <= nblk; ++r)
    for (int c = 1; c <= nblk; ++c)
      if (lcg_uniform() < occ && n < cap) {
        l[3 * n] = r;
        l[3 * n + 1] = c;
        l[3 * n + 2] = 1 + n * blk_elems;
        ++n;
      }
  *list3_out = l;
  return n;
}

int main(int argc, char** argv) {
  if (argc < 8) {
    fprintf(stderr, "usage: kbench <library.so> <out_dir> <nblk> <occ> <steps> <bsz> <variant:balance:chunk[:t]> ...\n");
    return 1;
  }
  const char* out_dir = argv[2];
  const int nblk = atoi(argv[3]);
  const double occ = atof(argv[4]);
  const int steps = atoi(argv[5]);
  int bm = 0, bn = 0, bk = 0;
  if (sscanf(argv[6], "%d,%d,%d", &bm, &bn, &bk) != 3) bm = bn = bk = atoi(argv[6]);
  const int bsz = bm; /* label of the result line for cubic shapes */
  const int first_spec = 7;
  const double t_start = now();
  lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (lib == NULL) {
    fprintf(stderr, "kbench: dlopen %s: %s\n", argv[1], dlerror());
    return 2;
  }
  acc_lib = lib;
  const char* other = getenv("KBENCH_ACC_LIB");
  if (other != NULL && other[0] != 0) {
    acc_lib = dlopen(other, RTLD_NOW | RTLD_LOCAL | RTLD_DEEPBIND); /* its internal calls must bind to its own definitions */
    if (acc_lib == NULL) {
      fprintf(stderr, "kbench: dlopen %s: %s\n", other, dlerror());
      return 2;
    }
    printf("kbench: acc/libsmm ABI from %s\n", other);
  }
  fn_i_v acc_init = (fn_i_v)sym_from(acc_lib, "c_dbcsr_acc_init");
  fn_set_dev set_dev = (fn_set_dev)sym_from(acc_lib, "c_dbcsr_acc_set_active_device");
  fn_stream_create stream_create = (fn_stream_create)sym_from(acc_lib, "c_dbcsr_acc_stream_create");
  fn_stream_sync stream_sync = (fn_stream_sync)sym_from(acc_lib, "c_dbcsr_acc_stream_sync");
  fn_dev_alloc dev_alloc = (fn_dev_alloc)sym_from(acc_lib, "c_dbcsr_acc_dev_mem_allocate");
  fn_dev_free dev_free = (fn_dev_free)sym_from(acc_lib, "c_dbcsr_acc_dev_mem_deallocate");
  fn_memcpy h2d = (fn_memcpy)sym_from(acc_lib, "c_dbcsr_acc_memcpy_h2d");
  fn_memcpy d2h = (fn_memcpy)sym_from(acc_lib, "c_dbcsr_acc_memcpy_d2h");
  fn_memset memset_zero = (fn_memset)sym_from(acc_lib, "c_dbcsr_acc_memset_zero");
  fn_process process = (fn_process)sym_from(acc_lib, "libsmm_acc_process");
  fn_norms block_norms = (fn_norms)sym("libsmm_acc_b200_block_norms_f64");
  fn_set_tun set_tun = acc_lib == lib ? (fn_set_tun)sym("libsmm_acc_b200_set_tunable") : noop_set_tun;
  fn_get_tun get_tun = (fn_get_tun)sym("libsmm_acc_b200_get_tunable");
  fn_set_trace set_trace = (fn_set_trace)sym("libsmm_acc_b200_set_trace");
  fn_cfg_default cfg_default = (fn_cfg_default)sym("dbcsr_b200_cfg_default");
  fn_eng_create eng_create = (fn_eng_create)sym("dbcsr_b200_engine_create");
  fn_eng_multiply eng_multiply = (fn_eng_multiply)sym("dbcsr_b200_engine_multiply");
  fn_eng_int eng_nstacks = (fn_eng_int)sym("dbcsr_b200_engine_nstacks");
  fn_eng_int_t eng_c_nblks = (fn_eng_int_t)sym("dbcsr_b200_engine_c_nblks");
  fn_eng_int_t eng_c_datasize = (fn_eng_int_t)sym("dbcsr_b200_engine_c_datasize");
  fn_eng_ptr_t eng_c_blk_p = (fn_eng_ptr_t)sym("dbcsr_b200_engine_c_blk_p");
  fn_eng_ptr_t eng_stack_dev = (fn_eng_ptr_t)sym("dbcsr_b200_engine_stack_dev");
  fn_eng_info eng_stack_info = (fn_eng_info)sym("dbcsr_b200_engine_stack_info");
  fn_eng_flop eng_flop = (fn_eng_flop)sym("dbcsr_b200_engine_flop");
  fn_eng_destroy eng_destroy = (fn_eng_destroy)sym("dbcsr_b200_engine_destroy");

  /* ---- workload: patterns, stacks (host only) */
  int *a_list, *b_list;
  const int na = make_pattern(nblk, occ, bm * bk, &a_list), nb = make_pattern(nblk, occ, bk * bn, &b_list);
  int* sizes = (int*)malloc(sizeof(int) * (size_t)nblk);
  int* sizes_n = (int*)malloc(sizeof(int) * (size_t)nblk);
  int* sizes_k = (int*)malloc(sizeof(int) * (size_t)nblk);
  for (int i = 0; i < nblk; ++i) sizes[i] = bm, sizes_n[i] = bn, sizes_k[i] = bk;
  dbcsr_b200_cfg_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg_default(&cfg);
  dbcsr_b200_engine_t* eng = eng_create(&cfg, sizes, nblk, sizes_n, nblk, sizes_k, nblk, 1, DBCSR_B200_RECORD, 0);
  if (eng == NULL) {
    fprintf(stderr, "kbench: engine_create failed\n");
    return 3;
  }
  CHECK(eng_multiply(eng, a_list, na, NULL, b_list, nb, NULL));
  const int nstacks = eng_nstacks(eng);
  const long long flop = eng_flop(eng);
  const int c_nblks = eng_c_nblks(eng, 0);
  const size_t c_elems = (size_t)eng_c_datasize(eng, 0);
  size_t total_entries = 0;
  int* st_size = (int*)malloc(sizeof(int) * (size_t)nstacks);
  size_t* st_off = (size_t*)malloc(sizeof(size_t) * (size_t)(nstacks + 1));
  for (int i = 0; i < nstacks; ++i) {
    int info[10];
    eng_stack_info(eng, i, info);
    if (info[0] != bm || info[1] != bn || info[2] != bk || info[6] != 1) {
      fprintf(stderr, "kbench: unexpected stack shape\n");
      return 3;
    }
    st_size[i] = info[7];
    st_off[i] = total_entries;
    total_entries += (size_t)info[7];
  }
  st_off[nstacks] = total_entries;
  int* all_stacks = (int*)malloc(sizeof(int) * 3 * total_entries);
  for (int i = 0; i < nstacks; ++i) memcpy(all_stacks + 3 * st_off[i], eng_stack_dev(eng, i), sizeof(int) * 3 * (size_t)st_size[i]);
  /* C block offsets (0-based) and sizes for the per-block norms */
  int* c_off = (int*)malloc(sizeof(int) * (size_t)c_nblks);
  int* c_len = (int*)malloc(sizeof(int) * (size_t)c_nblks);
  const int* blk_p = eng_c_blk_p(eng, 0);
  for (int i = 0; i < c_nblks; ++i) {
    c_off[i] = blk_p[i] - 1;
    c_len[i] = bm * bn;
  }
  const double t_built = now();
  printf("kbench: nblk %d occ %.3f  A %d B %d blocks  products %zu  stacks %d  C %d blocks  flop %lld  (host setup %.2f s)\n", nblk, occ,
         na, nb, total_entries, nstacks, c_nblks, flop, t_built - t_start);
  fflush(stdout);

  /* ---- device */
  CHECK(set_dev(0));
  CHECK(acc_init());
  void* stream = NULL;
  CHECK(stream_create(&stream, "kbench", 0));
  /* the sweep stream carries nothing but stack drains between synchronisations: declare it a chain (programmatic dependent launch
   * without the grid-dependency wait in front of the reads); KBENCH_NO_CHAIN=1 measures the waiting mode instead */
  {
    typedef int (*fn_chain)(void*, int);
    fn_chain chain = acc_lib == lib ? (fn_chain)dlsym(lib, "libsmm_acc_b200_stream_chain") : NULL;
    const char* nc = getenv("KBENCH_NO_CHAIN");
    const int on = chain != NULL && !(nc != NULL && nc[0] == '1');
    if (on) CHECK(chain(stream, 1));
    printf("kbench: chain mode %s\n", on ? "on" : "off");
  }
  const size_t a_elems = (size_t)na * bm * bk, b_elems = (size_t)nb * bk * bn;
  const size_t ab_max = a_elems > b_elems ? a_elems : b_elems;
  double* h = (double*)malloc(sizeof(double) * ab_max);
  for (size_t i = 0; i < ab_max; ++i) h[i] = A_VAL(i);
  void *d_a, *d_b, *d_c, *d_st, *d_off, *d_len, *d_norm;
  CHECK(dev_alloc(&d_a, a_elems * 8));
  CHECK(dev_alloc(&d_b, b_elems * 8));
  CHECK(dev_alloc(&d_c, c_elems * 8));
  CHECK(dev_alloc(&d_st, total_entries * 12));
  CHECK(dev_alloc(&d_off, (size_t)c_nblks * 4));
  CHECK(dev_alloc(&d_len, (size_t)c_nblks * 4));
  CHECK(dev_alloc(&d_norm, (size_t)c_nblks * 8));
  CHECK(h2d(h, d_a, a_elems * 8, stream));
  CHECK(stream_sync(stream));
  for (size_t i = 0; i < b_elems; ++i) h[i] = B_VAL(i);
  CHECK(h2d(h, d_b, b_elems * 8, stream));
  CHECK(h2d(all_stacks, d_st, total_entries * 12, stream));
  CHECK(h2d(c_off, d_off, (size_t)c_nblks * 4, stream));
  CHECK(h2d(c_len, d_len, (size_t)c_nblks * 4, stream));
  CHECK(stream_sync(stream));
  free(h);
  const size_t trace_words = (size_t)3 * 4096 * 128;
  void* d_trace;
  CHECK(dev_alloc(&d_trace, trace_words * 8));
  double* norms0 = (double*)malloc(sizeof(double) * (size_t)c_nblks);
  double* norms = (double*)malloc(sizeof(double) * (size_t)c_nblks);
  unsigned long long* h_trace = (unsigned long long*)malloc(trace_words * 8);
  printf("kbench: device ready (%.2f s), experiment build: %lld\n", now() - t_built, get_tun("experiment"));
  fflush(stdout);

  char path[1024];
  snprintf(path, sizeof(path), "%s/kbench_results.txt", out_dir);
  FILE* res = fopen(path, "a");

  for (int s = first_spec; s < argc; ++s) {
    int variant = 0, balance = 0, chunk = 0;
    char tflag = 0;
    const int nf = sscanf(argv[s], "%d:%d:%d:%c", &variant, &balance, &chunk, &tflag);
    if (nf < 3) {
      fprintf(stderr, "kbench: bad spec %s\n", argv[s]);
      continue;
    }
    set_tun("variant", variant);
    if (balance < 0) {  /* negative: the library's per-shape policy decides about run alignment (and chunk < 0: about the chunk size) */
      set_tun("balance", 0);
      set_tun("align", -1);
    }
    else {
      set_tun("balance", balance & 1);
      set_tun("align", (balance >> 1) & 1);
    }
    set_tun("chunk", chunk);
    set_trace(NULL);
    /* parity run */
    CHECK(memset_zero(d_c, 0, c_elems * 8, stream));
    int rc_bad = 0;
    for (int i = 0; i < nstacks; ++i) {
      const int rc = process(NULL, (const int*)d_st + 3 * st_off[i], st_size[i], 3, d_a, d_b, d_c, bm, bn, bk, 80, 1, stream, stream);
      if (rc != 0) rc_bad = rc;
    }
    CHECK(block_norms((const double*)d_c, c_nblks, (const int*)d_off, (const int*)d_len, (double*)d_norm, stream));
    CHECK(d2h(d_norm, norms, (size_t)c_nblks * 8, stream));
    CHECK(stream_sync(stream));
    long long mismatches = 0;
    double total = 0.0;
    for (int i = 0; i < c_nblks; ++i) total += norms[i];
    if (s == first_spec) {
      memcpy(norms0, norms, sizeof(double) * (size_t)c_nblks);
      /* absolute check of the reference spec: a few C blocks recomputed on the host (exact: small integers) */
      const int probe[4] = {0, 1, c_nblks / 2, c_nblks - 1};
      for (int q = 0; q < 4; ++q) {
        const int cf = c_off[probe[q]] + 1;
        double* blk = (double*)calloc((size_t)bm * bn, sizeof(double));
        for (size_t e = 0; e < total_entries; ++e)
          if (all_stacks[3 * e + 2] == cf) {
            const size_t a0 = (size_t)all_stacks[3 * e] - 1, b0 = (size_t)all_stacks[3 * e + 1] - 1;
            for (int j = 0; j < bn; ++j)
              for (int i = 0; i < bm; ++i) {
                double acc = 0.0;
                for (int k = 0; k < bk; ++k) acc += A_VAL(a0 + i + (size_t)k * bm) * B_VAL(b0 + j + (size_t)k * bn);
                blk[i + j * bm] += acc;
              }
          }
        double n2 = 0.0;
        for (int i = 0; i < bm * bn; ++i) n2 += blk[i] * blk[i];
        free(blk);
        printf("kbench: host check of C block %d: %s (host %.17g, device %.17g)\n", probe[q], n2 == norms[probe[q]] ? "exact" : "MISMATCH", n2,
               norms[probe[q]]);
      }
    }
    else
      for (int i = 0; i < c_nblks; ++i) mismatches += (norms[i] != norms0[i]);
    /* timing */
    double best = 1e30, sum = 0.0;
    const char* nosync = getenv("KBENCH_NOSYNC");
    if (nosync != NULL && nosync[0] == '1') {
      /* all repetitions enqueued back to back, ONE synchronisation at the end (deep launch queue, like a multi-step bench loop) */
      for (int i = 0; i < nstacks; ++i)
        process(NULL, (const int*)d_st + 3 * st_off[i], st_size[i], 3, d_a, d_b, d_c, bm, bn, bk, 80, 1, stream, stream);
      CHECK(stream_sync(stream));
      const double t0 = now();
      for (int it = 0; it < steps; ++it)
        for (int i = 0; i < nstacks; ++i)
          process(NULL, (const int*)d_st + 3 * st_off[i], st_size[i], 3, d_a, d_b, d_c, bm, bn, bk, 80, 1, stream, stream);
      const double t_enq = now() - t0;
      CHECK(stream_sync(stream));
      sum = now() - t0;
      best = sum / steps;
      printf("kbench: no-sync mode: %d drains enqueued in %.3f ms (%.2f us per launch on the host), done after %.3f ms\n", steps, t_enq * 1e3,
             t_enq * 1e6 / ((double)steps * nstacks), sum * 1e3);
    }
    else
    for (int it = 0; it < steps + 1; ++it) {
      CHECK(stream_sync(stream));
      const double t0 = now();
      for (int i = 0; i < nstacks; ++i)
        process(NULL, (const int*)d_st + 3 * st_off[i], st_size[i], 3, d_a, d_b, d_c, bm, bn, bk, 80, 1, stream, stream);
      CHECK(stream_sync(stream));
      const double dt = now() - t0;
      if (it > 0) {  /* first repetition = warm-up */
        sum += dt;
        if (dt < best) best = dt;
      }
    }
    const double mean = sum / steps;
    char line[512];
    snprintf(line, sizeof(line),
             "bsz %2d mnk %d,%d,%d spec %-10s variant %2d balance %d chunk %2d  rc %d  parity %s (mismatching blocks %lld, sum %.6e)  mean %.3f ms  best %.3f ms  "
             "%.2f TFLOP/s (best %.2f)  %.2f us/launch",
             bsz, bm, bn, bk, argv[s], variant, balance, chunk, rc_bad, (s == first_spec) ? "reference" : (mismatches == 0 ? "exact" : "MISMATCH"), mismatches, total,
             mean * 1e3, best * 1e3, flop / mean * 1e-12, flop / best * 1e-12, mean * 1e6 / nstacks);
    printf("%s\n", line);
    fflush(stdout);
    if (res != NULL) {
      fprintf(res, "%s\n", line);
      fflush(res);
    }
    if (nf == 4 && tflag == 't') {
      /* timeline of launches 100..102 of one more drain */
      CHECK(memset_zero(d_trace, 0, trace_words * 8, stream));
      CHECK(stream_sync(stream));
      set_tun("seq", 0);
      set_tun("trace_first", nstacks > 110 ? 100 : 0);
      set_tun("trace_count", 3);
      set_trace(d_trace);
      for (int i = 0; i < nstacks; ++i)
        process(NULL, (const int*)d_st + 3 * st_off[i], st_size[i], 3, d_a, d_b, d_c, bm, bn, bk, 80, 1, stream, stream);
      CHECK(stream_sync(stream));
      set_trace(NULL);
      CHECK(d2h(d_trace, h_trace, trace_words * 8, stream));
      CHECK(stream_sync(stream));
      snprintf(path, sizeof(path), "%s/trace_m%d_v%d_b%d_c%d.bin", out_dir, bsz, variant, balance, chunk);
      FILE* f = fopen(path, "wb");
      if (f != NULL) {
        fwrite(h_trace, 8, trace_words, f);
        fclose(f);
        printf("kbench: wrote %s\n", path);
      }
    }
  }
  if (res != NULL) fclose(res);
  printf("kbench: total %.1f s\n", now() - t_start);
  dev_free(d_trace);
  dev_free(d_a);
  dev_free(d_b);
  dev_free(d_c);
  dev_free(d_st);
  eng_destroy(eng);
  return 0;
}
