/*
 * include/dbcsr_b200_host.h -- C ABI of the host side of the local multiply (stack builder + accelerator driver).
 *
 * In a real DBCSR build this layer is DBCSR's own Fortran (it stays in place and only sees include/dbcsr_acc*.h):
 *   multrec   src/mm/dbcsr_mm_multrec.F:263-658       csr stack builder  src/mm/dbcsr_mm_csr.F:178-795
 *   scheduler src/mm/dbcsr_mm_sched.F:266-382          acc driver         src/mm/dbcsr_mm_accdrv.F:170-541
 *   index sort src/mm/dbcsr_mm_common.F:227-309        transposes         src/mm/dbcsr_mm_common.F:346-496
 * The functions below are the C++ stand-in used by bench.py / the tests (no Fortran compiler on the target boxes) and the
 * multi-threaded builder of SURVEY.md 8(f) row 1.  They call the accelerator ONLY through the acc/libsmm C ABI above.
 * Conventions as in DBCSR: block rows/cols and element offsets are 1-based; list index = (row, col, blk_p) triples.
 */
#ifndef DBCSR_B200_HOST_H
#define DBCSR_B200_HOST_H

#include <stddef.h>

#if defined(__cplusplus)
extern "C" {
#endif

/* dbcsr_cfg knobs on this path (src/core/dbcsr_config.F:151-180); zero-initialise then call dbcsr_b200_cfg_default */
typedef struct dbcsr_b200_cfg {
  int mm_stack_size;   /* MM_STACK_SIZE, 30000 on accelerator builds */
  int n_stacks;        /* N_STACKS, 3 */
  int multrec_limit;   /* MULTREC_LIMIT, 512 */
  int stack_sort;      /* ACCDRV_STACK_SORT, 1 */
  int min_flop_sort;   /* ACCDRV_MIN_FLOP_SORT, 4000 */
  int binning_nbins;   /* ACCDRV_BINNING_NBINS, 4096 */
  int binning_binsize; /* ACCDRV_BINNING_BINSIZE, 16 */
  int thread_buffers;  /* ACCDRV_THREAD_BUFFERS, 8 */
  int row_chunks;      /* (engine) block-row chunks per host thread, processed in row order; 1 = DBCSR's one slice per thread.
                          >1 lets the D2H of finished chunks start while later chunks are still being built/multiplied */
  int dev_tile;        /* (engine, DBCSR_B200_DEVICE_BUILD) 0 = stacks identical to the host builder's (DBCSR's traversal order).
                          T > 0 = tile order: inside every (row slice, stack number) group the products are ordered by T x T squares of
                          C blocks and by c_first inside a square before they are cut into stacks.  Same C index, same set of products;
                          every C block is accumulated in one run and a square's operands stay L2 resident (DRAM traffic of the stack
                          kernels drops several times); the order of summation differs from the reference's (results agree to rounding). */
} dbcsr_b200_cfg_t;
void dbcsr_b200_cfg_default(dbcsr_b200_cfg_t* cfg);

/* rec_sort_index (src/mm/dbcsr_mm_common.F:227-309), in place on nblks (row,col,blk_p) triples */
void dbcsr_b200_rec_sort_index(int nrows, int ncols, int nblks, int* list3);
/* same order; the top `depth` levels of the recursion sort their halves concurrently (what the engine uses for the right panel) */
void dbcsr_b200_rec_sort_index_mt(int nrows, int ncols, int nblks, int* list3, int depth);
/* stack_sort / stack_binning (src/mm/dbcsr_mm_accdrv.F:364-423): params7 -> out3 */
void dbcsr_b200_stack_sort(const int* params7, int* out3, int stack_size);
void dbcsr_b200_stack_binning(const int* params7, int* out3, int stack_size, int nbins, int binsize);

/* ---- engine: multrec + csr + sched + accdrv of `nthreads` host threads for one rank --------------------------------- */
typedef struct dbcsr_b200_engine dbcsr_b200_engine_t;

/* mode bits */
#define DBCSR_B200_LAUNCH 1 /* enqueue every dispatched stack on the accelerator (needs a device) */
#define DBCSR_B200_RECORD 2 /* keep every dispatched stack (host 7-wide + device-order 3-wide) for inspection / replay */
#define DBCSR_B200_DEVICE_BUILD 4 /* build the stacks and the C index ON THE DEVICE: the host only sorts the
                                     lists and walks the recursion of sparse_multrec down to its leaves; products, C blocks in first-touch
                                     order, stack filling / flushing and stack_sort run as data-parallel passes, with stacks, dispatch
                                     order and C index identical to the host builder's -- incl. existing C blocks (preset_c),
                                     retain_sparsity, symmetry skipping and the on-the-fly filter.  Engines with n_stacks^3 + 1 > 254
                                     use the host builder.  Without LAUNCH the same passes run in host loops (test harness on machines
                                     without a GPU). */

/* m_sizes/n_sizes/k_sizes: block sizes of the local C rows, C cols and the contraction index.
 * c_capacity: initial elements of every thread's device C buffer (0 = dense upper bound of the thread's block rows when that fits
 * comfortably into device memory, else an estimate from the panels); the buffer grows on demand.  Returns NULL on failure. */
dbcsr_b200_engine_t* dbcsr_b200_engine_create(const dbcsr_b200_cfg_t* cfg, const int* m_sizes, int nrows, const int* n_sizes,
  int ncols, const int* k_sizes, int nk, int nthreads, int mode, size_t c_capacity);
void dbcsr_b200_engine_destroy(dbcsr_b200_engine_t* e);

/* One Cannon tick (src/mm/dbcsr_mm_cannon.F:1622-1667 -> dbcsr_mm_multrec_multiply): a_list3/b_list3 are the panels' list
 * indices in BCSR order with panel-local coordinates; a_dev/b_dev the device data areas (B blocks already transposed, see
 * dbcsr_b200_transpose_panel).  Sorts the lists (rec_sort_index), splits the left list over the threads by block rows,
 * builds the stacks and (LAUNCH) streams them to the device.  Returns 0, or a negative code of the failing call. */
int dbcsr_b200_engine_multiply(dbcsr_b200_engine_t* e, const int* a_list3, int na, const void* a_dev, const int* b_list3, int nb,
  const void* b_dev);
/* On-the-fly norm filter (filter_eps of dbcsr_multiply; src/mm/dbcsr_mm_cannon.F:1038-1107, src/mm/dbcsr_mm_csr.F:270-278).
 * dbcsr_b200_row_max_epss: row_max_epss(r) = (filter_eps / max(1, total_row_counts(r)))^2 in single precision, where
 *   total_row_counts(r) = number of A blocks in block row r over the whole process row.
 * dbcsr_b200_engine_set_filter: thresholds for the engine's nrows local block rows (NULL switches the filter off).
 * dbcsr_b200_engine_multiply_filtered: like _multiply, with the per-block norms (sum of squares, single precision; what
 *   c_calculate_norms writes) of the panels, aligned with a_list3 / b_list3; a product is skipped when
 *   a_norm * b_norm < row_max_epss(a_row).  Without a preceding set_filter it behaves like _multiply. */
void dbcsr_b200_row_max_epss(double filter_eps, const int* total_row_counts, int nrows, float* row_max_epss);
int dbcsr_b200_engine_set_filter(dbcsr_b200_engine_t* e, const float* row_max_epss);
int dbcsr_b200_engine_multiply_filtered(dbcsr_b200_engine_t* e, const int* a_list3, int na, const void* a_dev, const float* a_norms,
  const int* b_list3, int nb, const void* b_dev, const float* b_norms);
/* Pipelined panel upload: the left panel can be uploaded in block-row pieces while earlier rows are already being multiplied and
 * their C blocks downloaded (PCIe is full duplex).  The engine's static row ownership: chunk c = block rows (row_lo, row_hi]
 * (1-based, dbcsr_b200_engine_chunk_rows), c < nchunks = nthreads * row_chunks, processed by thread c mod nthreads in chunk order.
 * events[c] (acc events, NULL = no wait): recorded by the caller behind the upload of the A blocks of chunk c; the next
 * dbcsr_b200_engine_multiply orders the stacks of chunk c behind events[c] and then forgets the list (single-tick multiplies). */
int dbcsr_b200_engine_nchunks(const dbcsr_b200_engine_t* e);
int dbcsr_b200_engine_chunk_rows(const dbcsr_b200_engine_t* e, int chunk, int* row_lo, int* row_hi);
int dbcsr_b200_engine_set_chunk_events(dbcsr_b200_engine_t* e, void* const* events, int nevents);
/* beta != 0 / retain_sparsity flows of dbcsr_multiply (src/mm/dbcsr_mm.F:706-709 scales C by beta first; the work matrices then start
 * from the existing blocks, src/mm/dbcsr_mm_csr.F:526-576).  rows/cols: block coordinates (1-based) of the existing C blocks;
 * host_data: their elements (col-major blocks, concatenated in list order, ALREADY scaled by beta) or NULL for zeros.
 * Block i is placed in the work area of the thread owning its block row, in list order; with LAUNCH its data is uploaded so that
 * the stack kernels accumulate onto it (the reference keeps a zeroed device buffer and block_adds on the host afterwards,
 * src/mm/dbcsr_mm_accdrv.F:340-362).  keep_sparsity != 0: products whose C block is not in the list are skipped
 * (src/mm/dbcsr_mm_csr.F:307).  Call after create/reset and before the first tick.  Both settings end with the next reset. */
int dbcsr_b200_engine_preset_c(dbcsr_b200_engine_t* e, const int* rows, const int* cols, int nblks, const double* host_data,
  int keep_sparsity);
/* Product matrix with symmetry (dbcsr_multiply with a symmetric / antisymmetric C; src/mm/dbcsr_mm_csr.F:280-292): of every
 * off-diagonal pair {(r,c),(c,r)} only the block whose GLOBAL coordinates need no transpose under DBCSR's checkerboard rule
 * checker_tr(row, col) = (odd(row + col) == (col >= row)) (src/dist/dbcsr_dist_operations.F:65-75) is computed.
 * global_rows / global_cols: global block index of every local C row / col (NULL = identity).  Ends with the next reset. */
int dbcsr_b200_engine_set_c_symmetry(dbcsr_b200_engine_t* e, int on, const int* global_rows, const int* global_cols);
/* Final filter of the product, multrec_filtering (src/mm/dbcsr_mm_multrec.F:700-758), index part: block b is kept iff blk_p[b] != 0,
 * nelems[b] != 0 and norms2[b] (= sum of squares of its elements, double) >= filter_eps^2.  Kept entries are moved to the front of
 * rows/cols/blk_p in order (blk_p values unchanged).  Returns the number kept; *nze_after = their element count. */
int dbcsr_b200_filter_index(double filter_eps, const double* norms2, int nblks, int* rows, int* cols, int* blk_p, const int* nelems,
  long long* nze_after);
/* The same filter run on the device BEFORE the download (LAUNCH engines, after the last tick): block norms by
 * libsmm_acc_b200_block_norms_f64, index compaction on the host, surviving blocks gathered into a contiguous device area in index
 * order.  Afterwards the c_* accessors, c_dev and c_to_host(_async) describe the filtered product (blk_p = compact offsets);
 * the next multiply/reset returns to the work matrices.  Not applied by DBCSR when retain_sparsity is set. */
int dbcsr_b200_engine_filter_c(dbcsr_b200_engine_t* e, double filter_eps);
/* dbcsr_finalize (work/dbcsr_work_operations.F:749-958), index part: sorts the nblks work-index entries (rows, cols; order of
 * first touch) into BCSR order - rows ascending, columns ascending within a row - in place, returns perm (sorted position ->
 * original position) and blk_p_new (1-based offsets of the data area compacted in that order; nelems = elements per ORIGINAL
 * entry) and *nze.  Returns 0, -5 if a block appears twice. */
int dbcsr_b200_finalize_index(int nblks, int* rows, int* cols, const int* nelems, int* perm, int* blk_p_new, long long* nze);
/* dbcsr_finalize on the device (LAUNCH engines, after the last tick): like dbcsr_b200_engine_filter_c - the final filter is
 * applied when filter_eps >= 0 - and in addition every thread's blocks are put into BCSR order before they are gathered, so the
 * downloaded data area of a thread is the final data area of its block rows.  Accessors as after filter_c. */
int dbcsr_b200_engine_finalize_c(dbcsr_b200_engine_t* e, double filter_eps);
/* block sizes of the contraction index of the NEXT panels (Cannon ticks bring different k-slices); the stack map built at
 * creation (from the k_sizes given there: use the global right-matrix row block sizes) is kept */
int dbcsr_b200_engine_set_k_sizes(dbcsr_b200_engine_t* e, const int* k_sizes, int nk);
/* begin a new multiply on the same (pooled) buffers: clears the product index, zeroes the device C buffers asynchronously */
int dbcsr_b200_engine_reset(dbcsr_b200_engine_t* e);
/* wait for all enqueued stacks (dbcsr_mm_accdrv_barrier, src/mm/dbcsr_mm_accdrv.F:425-431) */
int dbcsr_b200_engine_sync(dbcsr_b200_engine_t* e);

/* product work matrices, one per thread (pre-finalize index in order of first touch) */
/* (thread, tick) pairs whose stacks were built by the device passes since the engine was created (DBCSR_B200_DEVICE_BUILD) */
long long dbcsr_b200_engine_device_built_ticks(const dbcsr_b200_engine_t* e);
int dbcsr_b200_engine_nthreads(const dbcsr_b200_engine_t* e);
int dbcsr_b200_engine_c_nblks(const dbcsr_b200_engine_t* e, int thread);
int dbcsr_b200_engine_c_datasize(const dbcsr_b200_engine_t* e, int thread);
const int* dbcsr_b200_engine_c_rows(const dbcsr_b200_engine_t* e, int thread);
const int* dbcsr_b200_engine_c_cols(const dbcsr_b200_engine_t* e, int thread);
const int* dbcsr_b200_engine_c_blk_p(const dbcsr_b200_engine_t* e, int thread);
void* dbcsr_b200_engine_c_dev(const dbcsr_b200_engine_t* e, int thread);
/* D2H of thread's C buffer (datasize elements) into host memory (dbcsr_mm_accdrv_finalize, src/mm/dbcsr_mm_accdrv.F:340-362) */
int dbcsr_b200_engine_c_to_host(dbcsr_b200_engine_t* e, int thread, double* host);
/* asynchronous variant: enqueued on the thread's stream behind its last stack; complete after dbcsr_b200_engine_sync */
int dbcsr_b200_engine_c_to_host_async(dbcsr_b200_engine_t* e, int thread, double* host);
/* Optional: (pinned) host target for thread's C buffer; when set, the D2H of the thread's datasize elements is enqueued by the
 * thread itself right behind its last stack of every multiply (single-tick multiplies only), overlapping other threads' work;
 * complete after dbcsr_b200_engine_sync.  capacity (elements) of the thread's device C buffer is known after the first multiply. */
int dbcsr_b200_engine_set_c_host(dbcsr_b200_engine_t* e, int thread, double* host);
size_t dbcsr_b200_engine_c_capacity(const dbcsr_b200_engine_t* e, int thread);
/* all thread streams wait for an acc event (e.g. panels uploaded + transposed on another stream) */
int dbcsr_b200_engine_wait_event(dbcsr_b200_engine_t* e, void* event);
long long dbcsr_b200_engine_flop(const dbcsr_b200_engine_t* e);
/* seconds the host threads spent building+ordering stacks / waiting for free stack buffers in the last multiply (max over threads) */
double dbcsr_b200_engine_build_seconds(const dbcsr_b200_engine_t* e);

/* Statistics of the scheduler (dbcsr_mm_sched, src/mm/dbcsr_mm_sched.F:266-382,392-505; the table DBCSR prints at finalize),
 * accumulated over every multiply of the engine and merged over its threads.  table: up to max_rows rows of 7 values
 * (m, n, k, stack entries processed on the accelerator, stacks, stacks that ran on an untuned kernel, flop), ordered by flop;
 * inhomogeneous stacks are booked under (0,0,0).  totals[3] = flop, entries, stacks.  Returns the number of distinct (m,n,k). */
int dbcsr_b200_engine_stats(const dbcsr_b200_engine_t* e, long long* table, int max_rows, long long* totals);

/* Host-driver route of the scheduler (dbcsr_mm_sched_process, src/mm/dbcsr_mm_sched.F:340-363): when libsmm_acc_process
 * refuses a stack (negative return code, C untouched: inhomogeneous stack with DBCSR_B200_INHOMOGENEOUS=0, unsupported type, ...)
 * DBCSR drains it with its CPU driver into the HOST work matrix and adds the downloaded device buffer to it at finalize
 * (src/mm/dbcsr_mm_accdrv.F:340-362).  This library has no CPU compute path; the route exists when the caller installs a driver:
 * fn(ctx, thread, m, n, k, defined_mnk, params7, stack_size, c_datasize) must apply the stack (7 ints per entry, 1-based offsets,
 * B blocks as they are on the device) to the caller's host work area of `thread` (at least c_datasize elements) and return 0.
 * Without a driver a refused stack fails the multiply with the accelerator's code.  stats_cpu: totals[3] = flop, entries, stacks
 * that took this route. */
typedef int (*dbcsr_b200_host_driver_fn)(void* ctx, int thread, int m, int n, int k, int defined_mnk, const int* params7, int stack_size,
  int c_datasize);
int dbcsr_b200_engine_set_host_driver(dbcsr_b200_engine_t* e, dbcsr_b200_host_driver_fn fn, void* ctx);
int dbcsr_b200_engine_stats_cpu(const dbcsr_b200_engine_t* e, long long* totals);

/* recorded stacks (RECORD mode), in dispatch order per thread then concatenated thread by thread */
int dbcsr_b200_engine_nstacks(const dbcsr_b200_engine_t* e);
/* info[10] = m, n, k, max_m, max_n, max_k, defined_mnk, stack_size, thread, stack_number */
void dbcsr_b200_engine_stack_info(const dbcsr_b200_engine_t* e, int i, int* info);
const int* dbcsr_b200_engine_stack_host(const dbcsr_b200_engine_t* e, int i); /* 7 ints per entry */
const int* dbcsr_b200_engine_stack_dev(const dbcsr_b200_engine_t* e, int i);  /* 3 ints per entry, accdrv order */

/* ---- replay of one rank's whole Cannon multiply on PRE-BUILT device stacks (the multi-GPU "stack-kernel only" measurement, cf.
 * src/acc/acc_bench.c:338-345; schedule from dbcsr_b200/cannon.py, src/mm/dbcsr_mm_cannon.F:839-1771).  Ticks are numbered in the
 * order the rank takes them.  Per tick: device-to-device copies that fetch the tick's panels from their home ranks (peer memory
 * mapped by the caller; copy engines, issued on an internal side stream and ordered by events), the panels' device pointers,
 * and the tick's stacks (device-order int32 triples resident in HBM).  set_c: one or two pooled C buffers of `bytes` bytes;
 * zero_overlap != 0 with two buffers zeroes the next step's buffer on a side stream while this step's stacks run, else the
 * buffer is zeroed in line.  replay_step enqueues a whole multiply on `compute_stream` (an acc stream handle) without any host
 * synchronisation and returns 0 or the negative code of the failing call; replay_current_c = the buffer the last step wrote. */
typedef struct dbcsr_b200_replay dbcsr_b200_replay_t;
dbcsr_b200_replay_t* dbcsr_b200_replay_create(int nticks);
void dbcsr_b200_replay_destroy(dbcsr_b200_replay_t* r);
int dbcsr_b200_replay_set_panels(dbcsr_b200_replay_t* r, int tick, const void* a_dev, const void* b_dev);
int dbcsr_b200_replay_add_pull(dbcsr_b200_replay_t* r, int tick, void* dst, const void* src, size_t bytes);
int dbcsr_b200_replay_add_stack(dbcsr_b200_replay_t* r, int tick, const int* dev_stack, int size, int m, int n, int k, int defined_mnk);
int dbcsr_b200_replay_set_c(dbcsr_b200_replay_t* r, void* c0, void* c1, size_t bytes, int zero_overlap);
void* dbcsr_b200_replay_current_c(const dbcsr_b200_replay_t* r);
int dbcsr_b200_replay_step(dbcsr_b200_replay_t* r, void* compute_stream);

/* acc_transpose_blocks (src/mm/dbcsr_mm_common.F:346-496): in-place transpose of every block of the right panel on the device,
 * one libsmm_acc_transpose call per distinct (k,n) size pair.  b_list3 = (row=k index, col, blk_p). Synchronous w.r.t. `stream`
 * ordering only (no host sync).  scratch_dev must hold nb ints. */
int dbcsr_b200_transpose_panel(const int* b_list3, int nb, const int* k_sizes, const int* n_sizes, void* b_dev, int* scratch_host,
  void* scratch_dev, void* stream);
/* The same with the blocks' squared norms (float, list order; what acc_calculate_norms delivers for the on-the-fly filter,
 * src/mm/dbcsr_mm_common.F:498-591) computed in the SAME pass over the panel (libsmm_acc_b200_transpose_norms): dev_norms holds nb floats,
 * scratch_host / scratch_dev 2*nb ints.  Returns -3 when the panel has a block with a dimension above max_kernel_dim (80): those are not
 * transposed, the panel is left as libsmm_acc_transpose leaves it and the caller falls back to c_calculate_norms. */
int dbcsr_b200_transpose_panel_norms(const int* b_list3, int nb, const int* k_sizes, const int* n_sizes, void* b_dev, int* scratch_host,
  void* scratch_dev, float* dev_norms, void* stream);

#if defined(__cplusplus)
}
#endif

#endif /* DBCSR_B200_HOST_H */
