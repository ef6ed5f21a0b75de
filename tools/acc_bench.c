/*
 * tools/acc_bench.c -- stand-alone C miniapp that drives ONLY the drop-in C ABI (include/dbcsr_acc.h, include/dbcsr_acc_libsmm.h),
 * modelled on the reference's miniapp src/acc/acc_bench.c (which needs the external LIBXS library and therefore cannot be built
 * here): CLI `acc_bench [nrepeat] [stack_size] [m] [n] [k] [nc] [na] [nb]`, pinned + device buffers, synthetic C-sorted stack
 * (same recipe as INIT_STACK, src/acc/acc_bench.h:48-79, rand()-driven branch, srand(25071975) like acc_bench.c:150),
 * H2D, libsmm_acc_transpose warm-up, libsmm_acc_process x nrepeat timed with a host timer around the stream sync
 * (acc_bench.c:338-351), GFLOPS/s = 2*m*n*k*stack_size*nrepeat / t, then validation against a host loop with the stack semantics
 * of the CPU path (B not transposed on the host, src/mm/dbcsr_mm_hostdrv.F:248-282).  Environment: DEVICE, CHECK (max rel. error).
 * Build: gcc -O2 -Iinclude -o tools/acc_bench tools/acc_bench.c -Ldbcsr_b200/lib -ldbcsr_acc_b200 -lm
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "dbcsr_acc.h"
#include "dbcsr_acc_libsmm.h"

#define MAX_KERNEL_DIM 80
#define CHK(call) \
  do { \
    const int rc_ = (call); \
    if (0 != rc_) { \
      fprintf(stderr, "ERROR: %s returned %i (line %i)\n", #call, rc_, __LINE__); \
      return EXIT_FAILURE; \
    } \
  } while (0)

static double now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char* argv[]) {
  const int nrepeat = (1 < argc ? atoi(argv[1]) : 5);
  const int stack_size = (2 < argc ? atoi(argv[2]) : 30000);
  const int m = (3 < argc ? atoi(argv[3]) : 23);
  const int n = (4 < argc ? atoi(argv[4]) : m);
  const int k = (5 < argc ? atoi(argv[5]) : m);
  const int nc = (6 < argc ? atoi(argv[6]) : (stack_size / 16 > 0 ? stack_size / 16 : 1));
  const int na = (7 < argc ? atoi(argv[7]) : 10 * nc);
  const int nb = (8 < argc ? atoi(argv[8]) : 10 * nc);
  const char* const env_device = getenv("DEVICE");
  const char* const env_check = getenv("CHECK");
  const double check = (NULL != env_check ? atof(env_check) : 1e-10);
  const int mn = m * n, mk = m * k, kn = k * n;
  int ndevices = 0, i, r;
  void* stream = NULL;
  double *a_hst = NULL, *b_hst = NULL, *c_hst = NULL, *a_dev = NULL, *b_dev = NULL, *c_dev = NULL, *gold = NULL;
  int *stack_hst = NULL, *stack_dev = NULL, *trans_hst = NULL, *trans_dev = NULL, *host7 = NULL;
  double t0, duration, h2d_gbs, gflops, maxdiff = 0, maxref = 0;

  CHK(c_dbcsr_acc_get_ndevices(&ndevices));
  if (0 >= ndevices) {
    fprintf(stderr, "ERROR: No ACC-device found!\n");
    return EXIT_FAILURE;
  }
  CHK(c_dbcsr_acc_set_active_device(NULL != env_device ? atoi(env_device) : 0));
  CHK(c_dbcsr_acc_init());
  CHK(libsmm_acc_init());
  printf("%s\n", libsmm_acc_b200_version());
  printf("typename (id=%i): double\n", (int)dbcsr_type_real_8);
  printf("%s %i %i %i %i %i %i %i %i\n", argv[0], nrepeat, stack_size, m, n, k, nc, na, nb);
  CHK(c_dbcsr_acc_stream_create(&stream, "stream", -1));
  CHK(c_dbcsr_acc_host_mem_allocate((void**)&a_hst, sizeof(double) * mk * na, stream));
  CHK(c_dbcsr_acc_host_mem_allocate((void**)&b_hst, sizeof(double) * kn * nb, stream));
  CHK(c_dbcsr_acc_host_mem_allocate((void**)&c_hst, sizeof(double) * mn * nc, stream));
  CHK(c_dbcsr_acc_host_mem_allocate((void**)&stack_hst, sizeof(int) * 3 * stack_size, stream));
  CHK(c_dbcsr_acc_host_mem_allocate((void**)&trans_hst, sizeof(int) * nb, stream));
  CHK(c_dbcsr_acc_dev_mem_allocate((void**)&a_dev, sizeof(double) * mk * na));
  CHK(c_dbcsr_acc_dev_mem_allocate((void**)&b_dev, sizeof(double) * kn * nb));
  CHK(c_dbcsr_acc_dev_mem_allocate((void**)&c_dev, sizeof(double) * mn * nc));
  CHK(c_dbcsr_acc_dev_mem_allocate((void**)&stack_dev, sizeof(int) * 3 * stack_size));
  CHK(c_dbcsr_acc_dev_mem_allocate((void**)&trans_dev, sizeof(int) * nb));
  host7 = (int*)malloc(sizeof(int) * 7 * stack_size);
  gold = (double*)calloc((size_t)mn * nc, sizeof(double));
  if (NULL == host7 || NULL == gold) return EXIT_FAILURE;

  /* deterministic operands in (0,1] (INIT_MAT idea, src/acc/acc_bench.h:30-41, scaled per block) */
  for (i = 0; i < na; ++i)
    for (r = 0; r < mk; ++r) a_hst[(size_t)i * mk + r] = (double)((r + 1) * ((i % 7) + 1) % 97 + 1) / 98.0;
  for (i = 0; i < nb; ++i)
    for (r = 0; r < kn; ++r) b_hst[(size_t)i * kn + r] = (double)((r + 3) * ((i % 5) + 2) % 89 + 1) / 90.0;
  /* synthetic C-sorted stack: runs of navg +- nimb entries per C block, A/B blocks drawn with rand() */
  {
    const int navg = stack_size / nc, nimb = (1 > navg - 4 ? 1 : navg - 4);
    int idx = 0, c = 0, ntop = 0;
    srand(25071975);
    while (idx < stack_size) {
      const int rnd = rand(), next = c + 1;
      ntop += navg + (rnd % (2 * nimb) - nimb);
      if (stack_size < ntop) ntop = stack_size;
      for (; idx < ntop; ++idx) {
        const int ia = rand() % na, ib = rand() % nb;
        stack_hst[3 * idx + 0] = ia * mk + 1;
        stack_hst[3 * idx + 1] = ib * kn + 1;
        stack_hst[3 * idx + 2] = c * mn + 1;
        host7[7 * idx + 0] = m, host7[7 * idx + 1] = n, host7[7 * idx + 2] = k;
        host7[7 * idx + 3] = ia * mk + 1, host7[7 * idx + 4] = ib * kn + 1, host7[7 * idx + 5] = c * mn + 1, host7[7 * idx + 6] = c + 1;
      }
      if (next < nc) c = next;
    }
  }
  for (i = 0; i < nb; ++i) trans_hst[i] = i * kn; /* 0-based offsets of the B blocks */

  t0 = now();
  CHK(c_dbcsr_acc_memcpy_h2d(a_hst, a_dev, sizeof(double) * mk * na, stream));
  CHK(c_dbcsr_acc_memcpy_h2d(b_hst, b_dev, sizeof(double) * kn * nb, stream));
  CHK(c_dbcsr_acc_memcpy_h2d(stack_hst, stack_dev, sizeof(int) * 3 * stack_size, stream));
  CHK(c_dbcsr_acc_memcpy_h2d(trans_hst, trans_dev, sizeof(int) * nb, stream));
  CHK(c_dbcsr_acc_stream_sync(stream));
  duration = now() - t0;
  h2d_gbs = (sizeof(double) * ((double)mk * na + (double)kn * nb) + sizeof(int) * (3.0 * stack_size + nb)) / duration * 1e-9;
  printf("copy-in (%i MB): %.2g ms %.1f GB/s\n", (int)((sizeof(double) * ((size_t)mk * na + (size_t)kn * nb)) >> 20), 1e3 * duration, h2d_gbs);

  /* right panel: in-place transpose on the device (k x n -> n x k), like acc_transpose_blocks */
  t0 = now();
  CHK(libsmm_acc_transpose(trans_dev, 0, nb, b_dev, dbcsr_type_real_8, k, n, MAX_KERNEL_DIM, stream));
  CHK(c_dbcsr_acc_stream_sync(stream));
  printf("transpose: %.2g ms\n", 1e3 * (now() - t0));

  /* warm-up + timed stack drains */
  CHK(c_dbcsr_acc_memset_zero(c_dev, 0, sizeof(double) * mn * nc, stream));
  r = libsmm_acc_process(host7, stack_dev, stack_size, dbcsr_type_real_8, a_dev, b_dev, c_dev, m, n, k, MAX_KERNEL_DIM, 1, stream, stream);
  if (0 > r) {
    fprintf(stderr, "ERROR: libsmm_acc_process returned %i\n", r);
    return EXIT_FAILURE;
  }
  CHK(c_dbcsr_acc_memset_zero(c_dev, 0, sizeof(double) * mn * nc, stream));
  CHK(c_dbcsr_acc_stream_sync(stream));
  t0 = now();
  for (i = 0; i < nrepeat; ++i)
    (void)libsmm_acc_process(host7, stack_dev, stack_size, dbcsr_type_real_8, a_dev, b_dev, c_dev, m, n, k, MAX_KERNEL_DIM, 1, stream, stream);
  CHK(c_dbcsr_acc_stream_sync(stream));
  duration = now() - t0;
  gflops = 2.0 * m * n * k * (double)stack_size * nrepeat / duration * 1e-9;
  printf("device: %.2g ms %.1f GFLOPS/s (kernel kind %i, return code %i)\n", 1e3 * duration / nrepeat, gflops,
    libsmm_acc_b200_kernel_kind(m, n, k, dbcsr_type_real_8), r);

  /* validation: host loop over the 7-wide stack with UNtransposed B (CPU-path semantics), nrepeat times */
  CHK(c_dbcsr_acc_memcpy_d2h(c_dev, c_hst, sizeof(double) * mn * nc, stream));
  CHK(c_dbcsr_acc_stream_sync(stream));
  for (i = 0; i < stack_size; ++i) {
    const double* const a = a_hst + host7[7 * i + 3] - 1;
    const double* const b = b_hst + host7[7 * i + 4] - 1;
    double* const c = gold + host7[7 * i + 5] - 1;
    int col, l, row;
    for (col = 0; col < n; ++col)
      for (l = 0; l < k; ++l) {
        const double t = b[col * k + l];
        for (row = 0; row < m; ++row) c[col * m + row] += t * a[l * m + row];
      }
  }
  for (i = 0; i < mn * nc; ++i) {
    const double ref = gold[i] * nrepeat, d = fabs(c_hst[i] - ref);
    if (maxdiff < d) maxdiff = d;
    if (maxref < fabs(ref)) maxref = fabs(ref);
  }
  printf("max.error: abs=%g rel=%g (limit %g)\n", maxdiff, maxdiff / (0 < maxref ? maxref : 1), check);

  CHK(c_dbcsr_acc_dev_mem_deallocate(a_dev));
  CHK(c_dbcsr_acc_dev_mem_deallocate(b_dev));
  CHK(c_dbcsr_acc_dev_mem_deallocate(c_dev));
  CHK(c_dbcsr_acc_dev_mem_deallocate(stack_dev));
  CHK(c_dbcsr_acc_dev_mem_deallocate(trans_dev));
  CHK(c_dbcsr_acc_host_mem_deallocate(a_hst, stream));
  CHK(c_dbcsr_acc_host_mem_deallocate(b_hst, stream));
  CHK(c_dbcsr_acc_host_mem_deallocate(c_hst, stream));
  CHK(c_dbcsr_acc_host_mem_deallocate(stack_hst, stream));
  CHK(c_dbcsr_acc_host_mem_deallocate(trans_hst, stream));
  CHK(c_dbcsr_acc_stream_destroy(stream));
  CHK(libsmm_acc_finalize());
  CHK(c_dbcsr_acc_finalize());
  free(host7);
  free(gold);
  if (maxdiff / (0 < maxref ? maxref : 1) > check) {
    fprintf(stderr, "FAILED\n");
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
