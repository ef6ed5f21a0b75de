/*
 * oracle/dbcsr_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the arithmetic on DBCSR's stack-drain hot path and of the
 * deterministic input generators the reference's own tests use.  Only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() may call this.
 * The product (libdbcsr_acc_b200.so) never links or loads it.
 *
 * Every function cites the reference file:line (relative to /root/reference) it restates.
 * Parity pinning: (1) tests/test_oracle_golden.py reproduces the eight golden checksums stored in the
 * reference's tests/inputs/ *.perf files through orc_dlarnv1 + orc_set_larnv_seed + block multiply +
 * orc_dbcsr_checksum; (2) tests/test_oracle_vs_ref.py compares orc_stack_calc / orc_mat_init /
 * orc_stack_init / orc_checksum / orc_transpose against the reference's own C++ checker functions
 * compiled from /root/reference into oracle/_ref/ (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(_OPENMP)
#  include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------
 * Device-stack semantics: C += A * Bt^T, B stored transposed (n x k col-major).
 * Restates stackCalc, src/acc/libsmm_acc/libsmm_acc_benchmark.cpp:126-145.
 * stack: 3 ints per entry (a_first, b_first, c_first), 1-based element offsets.
 * ---------------------------------------------------------------------------------------------- */
void orc_stack_calc(const int* stack, int n_stack, double* mat_c, const double* mat_a, const double* mat_b, int mat_m,
                    int mat_n, int mat_k) {
  for (int s = 0; s < n_stack; s++) {
    const int a_base = stack[3 * s] - 1;
    const int b_base = stack[3 * s + 1] - 1;
    const int c_base = stack[3 * s + 2] - 1;
    for (int n = 0; n < mat_n; n++) {
      for (int m = 0; m < mat_m; m++) {
        double res = 0.;
        for (int k = 0; k < mat_k; k++) res += mat_a[a_base + k * mat_m + m] * mat_b[b_base + k * mat_n + n];
        mat_c[c_base + n * mat_m + m] += res;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * Host-stack semantics (the reference CPU path): per entry DGEMM('N','N',m,n,k,1,A,m,B,k,1,C,m).
 * Restates blas_process_mm_stack_d, src/mm/dbcsr_mm_hostdrv.F:248-282.
 * params: 7 ints per entry (m,n,k,a_first,b_first,c_first,c_blk), 1-based offsets; B NOT transposed.
 * dgemm_fn: Fortran-ABI dgemm_ (e.g. scipy_dgemm_ from the OpenBLAS bundled with scipy) or NULL for the
 * textbook triple loop (reference BLAS loop order j,l,i).
 * ---------------------------------------------------------------------------------------------- */
typedef void (*dgemm_fn_t)(const char*, const char*, const int*, const int*, const int*, const double*, const double*,
                           const int*, const double*, const int*, const double*, double*, const int*);

static void naive_dgemm_nn(int m, int n, int k, const double* a, const double* b, double* c) {
  for (int j = 0; j < n; j++)
    for (int l = 0; l < k; l++) {
      const double t = b[j * k + l];
      for (int i = 0; i < m; i++) c[j * m + i] += t * a[l * m + i];
    }
}

void orc_host_stack(const int* params, int stack_size, const double* a_data, const double* b_data, double* c_data,
                    void* dgemm_fn) {
  const double one = 1.0;
  dgemm_fn_t f = (dgemm_fn_t)dgemm_fn;
  for (int sp = 0; sp < stack_size; sp++) {
    const int* p = params + 7 * sp;
    const int m = p[0], n = p[1], k = p[2];
    const double* a = a_data + (p[3] - 1);
    const double* b = b_data + (p[4] - 1);
    double* c = c_data + (p[5] - 1);
    if (f)
      f("N", "N", &m, &n, &k, &one, a, &m, b, &k, &one, c, &m);
    else
      naive_dgemm_nn(m, n, k, a, b, c);
  }
}

/* Threaded driver for the CPU baseline: stacks[i] are independent host stacks whose C blocks are disjoint
 * between stacks of different threads (DBCSR's model: every OpenMP thread owns disjoint C rows,
 * src/mm/dbcsr_mm_multrec.F:306-311).  stack_ptr[i]..stack_ptr[i+1] delimit stack i inside params.
 * owner[i] = thread that must run stack i (stacks of one owner run in order). */
void orc_host_stacks_threaded(const int* params, const long* stack_ptr, const int* owner, int n_stacks, int n_threads,
                              const double* a_data, const double* b_data, double* c_data, void* dgemm_fn) {
#if defined(_OPENMP)
#  pragma omp parallel num_threads(n_threads)
  {
    const int tid = omp_get_thread_num();
    for (int i = 0; i < n_stacks; i++)
      if (owner[i] == tid)
        orc_host_stack(params + 7 * stack_ptr[i], (int)(stack_ptr[i + 1] - stack_ptr[i]), a_data, b_data, c_data, dgemm_fn);
  }
#else
  (void)n_threads;
  (void)owner;
  for (int i = 0; i < n_stacks; i++)
    orc_host_stack(params + 7 * stack_ptr[i], (int)(stack_ptr[i + 1] - stack_ptr[i]), a_data, b_data, c_data, dgemm_fn);
#endif
}

int orc_max_threads(void) {
#if defined(_OPENMP)
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * In-place transpose of the m x n col-major blocks listed (0-based offsets) in stack.
 * Result equals transpose_d, src/acc/libsmm_acc/kernels/smm_acc_transpose.h:41-65 (out[i] = in[(i%n)*m + i/n]),
 * and stackTransp, src/acc/libsmm_acc/libsmm_acc_benchmark.cpp:148-161.
 * ---------------------------------------------------------------------------------------------- */
void orc_transpose(const int* stack, int n_stack, double* mat, int m, int n) {
  double* buf = (double*)malloc(sizeof(double) * (size_t)m * n);
  for (int s = 0; s < n_stack; s++) {
    double* blk = mat + stack[s];
    memcpy(buf, blk, sizeof(double) * (size_t)m * n);
    for (int i = 0; i < m * n; i++) blk[i] = buf[(i % n) * m + i / n];
  }
  free(buf);
}

/* Sum of squares per block, float out.  Restates calculate_norms_d, src/acc/cuda_hip/calculate_norms.cpp:48-96
 * (double accumulation, converted to float on store, no sqrt). */
void orc_norms(const double* mat, int nblks, const int* offsets, const int* nelems, float* norms) {
  for (int b = 0; b < nblks; b++) {
    double sum = 0.0;
    for (int i = 0; i < nelems[b]; i++) {
      const double d = mat[offsets[b] + i];
      sum += d * d;
    }
    norms[b] = (float)sum;
  }
}

/* matInit, src/acc/libsmm_acc/libsmm_acc_benchmark.cpp:103-109: integer-valued test data. */
void orc_mat_init(double* mat, int mat_n, int x, int y, int seed) {
  double* m = mat;
  for (int n = 0; n < mat_n; n++)
    for (int j = 0; j < y; j++)
      for (int i = 0; i < x; i++, m++) *m = (double)j * x + i + n + seed;
}

/* stackInit -> INIT_STACK, src/acc/acc_bench.h:48-79 (rand()-driven branch, rnd == NULL),
 * called from src/acc/libsmm_acc/libsmm_acc_benchmark.cpp:114-116.  Uses libc rand() exactly like the
 * reference (caller seeds with srand()). C-sorted synthetic stack: run lengths navg +- nimb. */
void orc_stack_init(int* stack, int stack_size, int nc, int na, int nb, int m, int n, int k) {
  const int mn = m * n, mk = m * k, kn = k * n;
  const int navg = stack_size / nc;
  const int nimb = (1 > navg - 4) ? 1 : navg - 4;
  int i = 0, c = 0, ntop = 0;
  int* p = stack;
  while (i < stack_size) {
    const int r = rand();
    const int next = c + 1;
    ntop += navg + (r % (2 * nimb) - nimb);
    if (stack_size < ntop) ntop = stack_size;
    for (; i < ntop; ++i) {
      const int a = rand() % na;
      const int b = rand() % nb;
      *p++ = a * mk + 1;
      *p++ = b * kn + 1;
      *p++ = c * mn + 1;
    }
    if (next < nc) c = next;
  }
}

/* checkSum, src/acc/libsmm_acc/libsmm_acc_benchmark.cpp:164-170. */
double orc_checksum(const double* mat_c, int n_c, int mat_m, int mat_n) {
  double res = 0;
  for (int i = 0; i < n_c * mat_m * mat_n; i++) res += mat_c[i];
  return res;
}

/* checkSumTransp, src/acc/libsmm_acc/libsmm_acc_benchmark.cpp:173-191. */
double orc_checksum_transp(const double* mat, int n_stack, int mat_m, int mat_n) {
  double res = 0;
  const int size = mat_m * mat_n;
  const int n_samples = size / 3;
  int step = size;
  if (n_samples > 0) step = size / n_samples;
  for (int s = 0; s < n_stack; s++) {
    const int offset = s * size;
    for (int idx = s % step; idx < size; idx += step) res += mat[offset + idx];
  }
  return res;
}

/* ------------------------------------------------------------------------------------------------
 * Random-matrix recipe of the reference's tests.
 * ---------------------------------------------------------------------------------------------- */

/* set_larnv_seed, src/utils/dbcsr_blas_operations.F:29-52. */
void orc_set_larnv_seed(int irow, int nrow, int icol, int ncol, int ival, int* iseed) {
  (void)ncol;
  int64_t map = (((int64_t)irow - 1 + (int64_t)icol * (int64_t)nrow) * (1 + (int64_t)(((ival % 65536) + 65536) % 65536))) * 2 + 1;
  iseed[3] = (int)(map % 4096);
  map /= 4096;
  iseed[2] = (int)((map ^ 3541) % 4096);
  map /= 4096;
  iseed[1] = (int)((map ^ 1153) % 4096);
  map /= 4096;
  iseed[0] = (int)((map ^ 2029) % 4096);
}

/* LAPACK DLARNV(IDIST=1) -> DLARUV (third-party: reference LAPACK 3.x, not vendored under /root/reference;
 * call sites src/ops/dbcsr_test_methods.F:399,423 via src/utils/dbcsr_blas_operations.F:54-80).
 * Published algorithm: multiplicative congruential generator x <- a*x mod 2^48, a = 33952834046453,
 * seed held as four 12-bit digits, uniform = x / 2^48 (exact in double).  DLARUV's table MM(i,:) is a^i,
 * so a call for n numbers is n LCG steps.  Pinned against scipy's OpenBLAS scipy_dlarnv_ in tests. */
void orc_dlarnv1(int* iseed, int n, double* x) {
  const uint64_t a = 33952834046453ULL, mask = (1ULL << 48) - 1;
  uint64_t s = ((uint64_t)iseed[0] << 36) | ((uint64_t)iseed[1] << 24) | ((uint64_t)iseed[2] << 12) | (uint64_t)iseed[3];
  for (int i = 0; i < n; i++) {
    s = (s * a) & mask;
    x[i] = (double)s * (1.0 / 281474976710656.0);
  }
  iseed[0] = (int)((s >> 36) & 4095);
  iseed[1] = (int)((s >> 24) & 4095);
  iseed[2] = (int)((s >> 12) & 4095);
  iseed[3] = (int)(s & 4095);
}

/* Block-presence pattern of dbcsr_make_random_matrix, src/ops/dbcsr_test_methods.F:392-410:
 * geometric skipping over the row-major numbering of the nrow x ncol block grid.
 * Writes 1-based (row, col) of present blocks (already in BCSR order); returns the count
 * (or -count-1 if cap was too small). counter = randmat_counter of that matrix (12341313 + call number). */
long orc_random_blocks(int nrow, int ncol, double sparsity, int counter, int* rows, int* cols, long cap) {
  int jseed[4];
  double value;
  const double my_sparsity = (sparsity > 1.0) ? sparsity / 100.0 : sparsity;
  const int64_t nmax = (int64_t)nrow * (int64_t)ncol;
  int64_t ele = -1;
  long count = 0;
  orc_set_larnv_seed(7, 42, 3, 42, counter, jseed);
  for (;;) {
    int64_t increment;
    orc_dlarnv1(jseed, 1, &value);
    if (my_sparsity > 0)
      increment = 1 + (int64_t)floor(log(value) / log(my_sparsity));
    else
      increment = 1;
    ele += increment;
    if (ele >= nmax) break;
    if (count < cap) {
      rows[count] = (int)(ele / ncol) + 1;
      cols[count] = (int)(ele % ncol) + 1;
    }
    count++;
  }
  return (count <= cap) ? count : -count - 1;
}

/* Block values of dbcsr_make_random_matrix, src/ops/dbcsr_test_methods.F:421-424:
 * uniform(0,1) from a per-block seed, col-major fill. */
void orc_fill_block(int row, int nrow, int col, int ncol, int counter, int nze, double* out) {
  int iseed[4];
  orc_set_larnv_seed(row, nrow, col, ncol, counter, iseed);
  orc_dlarnv1(iseed, nze, out);
}

/* Fill all blocks of a matrix (rows/cols 1-based, offsets 0-based element offsets into data). */
void orc_fill_blocks(long nblks, const int* rows, const int* cols, const long* offsets, const int* row_blk_size,
                     const int* col_blk_size, int nrow, int ncol, int counter, double* data) {
#if defined(_OPENMP)
#  pragma omp parallel for schedule(static)
#endif
  for (long b = 0; b < nblks; b++)
    orc_fill_block(rows[b], nrow, cols[b], ncol, counter, row_blk_size[rows[b] - 1] * col_blk_size[cols[b] - 1],
                   data + offsets[b]);
}

/* dbcsr_checksum, src/dist/dbcsr_dist_util.F:432-547 (+ pd_blk_cs :549-575) for real_8, untransposed blocks.
 * Blocks must be given in BCSR order (rows ascending); rows/cols 1-based; offsets 0-based;
 * row_off/col_off = 1-based first full row/col of each block row/col. pos != 0: position-dependent checksum. */
double orc_dbcsr_checksum(long nblks, const int* rows, const int* cols, const long* offsets, const int* row_blk_size,
                          const int* col_blk_size, const int* row_off, const int* col_off, const double* data, int pos) {
  double local_cs = 0.0;
  long b = 0;
  while (b < nblks) {
    const int br = rows[b];
    double local_cs_row = 0.0;
    for (; b < nblks && rows[b] == br; b++) {
      const int m = row_blk_size[br - 1], n = col_blk_size[cols[b] - 1];
      const double* d = data + offsets[b];
      double blk_cs = 0.0;
      if (pos) {
        const int ro = row_off[br - 1], co = col_off[cols[b] - 1];
        for (int c = 1; c <= n; c++)
          for (int r = 1; r <= m; r++)
            blk_cs += d[(c - 1) * m + (r - 1)] * log(fabs((double)(ro + r - 1) * (double)(co + c - 1)));
      }
      else {
        for (int i = 0; i < m * n; i++) blk_cs += d[i] * d[i];
      }
      local_cs_row += blk_cs;
    }
    local_cs += local_cs_row;
  }
  return local_cs;
}

/* Plain block product used by the golden-checksum test: C(m x n) += op(A) * B with col-major blocks.
 * ta != 0: A is stored k x m and used transposed (dbcsr_multiply transa='T', src/mm/dbcsr_mm.F:523-580). */
void orc_block_gemm(int m, int n, int k, const double* a, int ta, const double* b, double* c) {
  for (int j = 0; j < n; j++)
    for (int l = 0; l < k; l++) {
      const double t = b[j * k + l];
      if (!ta)
        for (int i = 0; i < m; i++) c[j * m + i] += t * a[l * m + i];
      else
        for (int i = 0; i < m; i++) c[j * m + i] += t * a[i * k + l];
    }
}
