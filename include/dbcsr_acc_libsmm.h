/*
 * include/dbcsr_acc_libsmm.h -- C ABI of the B200-native small-matrix-multiplication (SMM) backend.
 *
 * Drop-in boundary, part 2 of 2: the entry points DBCSR binds for the stack-drain hot path
 * (reference interface: src/acc/acc_libsmm.h:31-49; Fortran callers src/mm/dbcsr_acc_operations.F:38-67,120-131,171-179,
 * src/mm/dbcsr_mm_common.F:117-129, src/core/dbcsr_lib.F:86-116).
 */
#ifndef DBCSR_B200_ACC_LIBSMM_H
#define DBCSR_B200_ACC_LIBSMM_H

#include "dbcsr_acc.h"

#if defined(__cplusplus)
extern "C" {
#endif

/* acc_libsmm.h:31-36.  dbcsr_type_bf16_ext is an EXTENSION of this library (DBCSR has no 16-bit type): A/B panels
 * stored as bf16, C accumulated in fp32; same stack format, element offsets count elements of the respective type. */
typedef enum libsmm_acc_data_t {
  dbcsr_type_real_4 = 1,
  dbcsr_type_real_8 = 3,
  dbcsr_type_complex_4 = 5,
  dbcsr_type_complex_8 = 7,
  dbcsr_type_bf16_ext = 9
} libsmm_acc_data_t;

/* acc_libsmm.h:38-40.  init/finalize may be called more than once (src/core/dbcsr_lib.F:241).
 * is_thread_safe must report 1 for an OpenMP build of DBCSR (src/core/dbcsr_lib.F:248-261): this library is thread safe. */
int libsmm_acc_init(void);
int libsmm_acc_finalize(void);
c_dbcsr_acc_bool_t libsmm_acc_is_thread_safe(void);

/* acc_libsmm.h:42-43 (src/acc/libsmm_acc/libsmm_acc.cpp:482-487).  For i in [offset, offset+stack_size): in-place transpose
 * of the m x n column-major block at 0-BASED element offset dev_trs_stack[i] of dev_data (result n x m column-major).
 * Returns 0 and does nothing for datatype != real_8 or m,n > max_kernel_dim (reference behaviour); non-zero aborts DBCSR. */
int libsmm_acc_transpose(const int* dev_trs_stack, int offset, int stack_size, void* dev_data, libsmm_acc_data_t datatype, int m,
  int n, int max_kernel_dim, void* stream);

/* acc_libsmm.h:45-47 (src/acc/libsmm_acc/libsmm_acc.cpp:324-339).  Enqueue, without synchronising, for every entry i of the stack
 *     C[c_i .. c_i+m*n) += A[a_i .. a_i+m*k) (m x k col-major)  *  B_i ,   B_i = n x k col-major (i.e. stored TRANSPOSED)
 * when n,k <= max_kernel_dim, else k x n col-major.
 *   host_param_stack : 7 ints/entry (m,n,k,a_first,b_first,c_first,c_blk), 1-based, original order, host memory
 *   dev_param_stack  : 3 ints/entry (a_first,b_first,c_first), 1-based, (mostly) sorted by c_first, device memory,
 *                      its H2D copy already enqueued on stack_stream
 *   m_max,n_max,k_max: the (m,n,k) of every entry when def_mnk == 1
 * Return: 0 ok (specialised kernel), 10 ok (generic untuned kernel), <0 nothing was enqueued and C is untouched
 * (-1 inhomogeneous stack not handled, -10 datatype not handled): DBCSR then re-does the stack on the CPU
 * (src/mm/dbcsr_acc_operations.F:134-135). */
int libsmm_acc_process(const int* host_param_stack, const int* dev_param_stack, int stack_size, libsmm_acc_data_t datatype,
  const void* dev_a_data, const void* dev_b_data, void* dev_c_data, int m_max, int n_max, int k_max, int max_kernel_dim,
  c_dbcsr_acc_bool_t def_mnk, void* stack_stream, void* c_stream);

/* acc_libsmm.h:49 (src/acc/cuda_hip/calculate_norms.cpp:98-117): norms[b] = sum_i mat[offsets[b]+i]^2 (float, no sqrt),
 * offsets 0-based, all pointers device memory. */
int c_calculate_norms(const double* mat, int nblks, const int* offsets, const int* nelems, float* norms, void* stream_ptr);

/* Extensions for the product's finalize step on the device (the reference does both on the host after the D2H of C):
 * block_norms_f64: norms[b] = sum_i mat[offsets[b]+i]^2 in double - the quantity multrec_filtering compares with filter_eps^2
 *   (src/mm/dbcsr_mm_multrec.F:700-758: DDOT(blk,blk) >= filter_eps**2 keeps the block).  0-based offsets like c_calculate_norms.
 * gather_blocks: dst[dst_offsets[b]+i] = src[src_offsets[b]+i], i < nelems[b]; src and dst must not overlap. */
int libsmm_acc_b200_block_norms_f64(const double* mat, int nblks, const int* offsets, const int* nelems, double* norms, void* stream_ptr);
int libsmm_acc_b200_gather_blocks(const double* src, double* dst, int nblks, const int* src_offsets, const int* dst_offsets,
  const int* nelems, void* stream_ptr);

/* Transpose + block norms in ONE pass over the right panel (the reference runs libsmm_acc_transpose and, with
 * filter_eps, c_calculate_norms as two passes): libsmm_acc_transpose semantics for real_8, plus dev_norms[dev_trs_blk[i]] = sum of squares
 * of block i (float) for i in [offset, offset + stack_size); dev_trs_blk NULL = i.  -3 (nothing done) if m or n > max_kernel_dim. */
int libsmm_acc_b200_transpose_norms(const int* dev_trs_stack, const int* dev_trs_blk, int offset, int stack_size, double* dev_data, int m,
  int n, int max_kernel_dim, float* dev_norms, void* stream);

/* Declared by DBCSR (interface in src/core/dbcsr_lib.F:111-116) but never defined by the reference; exported for safety. */
int libsmm_acc_gpu_warp_size(void);

/* ---- extensions of this library (not part of the reference ABI) -------------------------------------------------- */

/* BF16 extension (dbcsr_type_bf16_ext): convert nblks FP64 blocks, stored back to back (block b at dev_src + b*rows*kdim, element
 * (row, kk) at [row*row_stride + kk*k_stride]), into BF16 operand tiles of libsmm_acc_b200_bf16_tile_bytes(rows, kdim) bytes each
 * (canonical tcgen05 K-major layout, zero padded).  A panel: rows = m, kdim = k, row_stride = 1, k_stride = m.  Transposed B panel
 * (n x k column-major, what libsmm_acc_transpose leaves behind): rows = n, kdim = k, row_stride = 1, k_stride = n.
 * libsmm_acc_process(datatype = dbcsr_type_bf16_ext) then takes the tile panels as dev_a_data / dev_b_data, an FP32 C buffer, and
 * the ordinary stack (element offsets of the ORIGINAL FP64 panels; all blocks of a panel must have the stack's (m,k) / (n,k)). */
int libsmm_acc_b200_pack_bf16(const double* dev_src, int nblks, int rows, int kdim, int row_stride, int k_stride, void* dev_dst,
  void* stream);
int libsmm_acc_b200_bf16_tile_bytes(int rows, int kdim);
/* Tiled BF16 SpGEMM (dbcsr_b200/csrc/smm_bf16_tiled.cuh; BASELINE config 4): for dense-ish products the multiply is driven by the
 * block index instead of parameter stacks.
 * pack_bf16_rk converts nblks FP64 blocks (addressing as in pack_bf16) into BF16 "rk" operand tiles (K padded to 32): the tile of block b
 *   goes to dev_dst + slot * dst_pitch with slot = dev_dst_slot[b] (NULL: b); dst_pitch = libsmm_acc_b200_bf16_rk_slot_bytes(rows,
 *   b_operand) (A operand: the tile size ceil(rows/8)*512; B operand: 2048), the gap behind a tile is zero-filled.
 *   Tiles of blocks that are adjacent in the operand should be adjacent in memory (the kernel then fetches them with one copy): A tiles in
 *   block-COLUMN order (slot = rank of the block in (k block, block row) order), B tiles in block-row order (= BCSR order, NULL).
 * bf16_spgemm computes, for every block row rb < nrb and block column cb < ncb with dev_c_off[rb*ncb + cb] >= 0, the FP32 block
 *   C(rb,cb) = sum over kb < nkb of A(rb,kb) * B(kb,cb) (m x n column-major at element offset dev_c_off[..]; blocks with no contribution are
 *   written as zeros; C is OVERWRITTEN, it need not be zeroed), where dev_a_map[kb*nrb + rb] / dev_b_map[kb*ncb + cb] is the SLOT of the
 *   packed tile of A(rb,kb) (rows = m, kdim = k) / of B(kb,cb) TRANSPOSED (rows = n, kdim = k) in a_tiles / b_tiles, or -1 for an absent
 *   block.  All block rows have m, all block columns n, all k blocks k elements; m, n, k <= 32 (else -10, nothing enqueued).
 *   Asynchronous on `stream`. */
int libsmm_acc_b200_bf16_rk_tile_bytes(int rows);
int libsmm_acc_b200_bf16_rk_slot_bytes(int rows, int b_operand);
int libsmm_acc_b200_pack_bf16_rk(const double* dev_src, int nblks, int rows, int kdim, int row_stride, int k_stride, void* dev_dst,
  int dst_pitch, const int* dev_dst_slot, void* stream);
int libsmm_acc_b200_bf16_spgemm(const void* a_tiles, const int* dev_a_map, const void* b_tiles, const int* dev_b_map, float* dev_c,
  const int* dev_c_off, int nrb, int ncb, int nkb, int m, int n, int k, void* stream);
/* Which kernel would libsmm_acc_process use for (m,n,k)?  0 = none, 1 = specialised DMMA kernel, 2 = generic kernel, 3 = BF16 tcgen05. */
int libsmm_acc_b200_kernel_kind(int m, int n, int k, libsmm_acc_data_t datatype);
/* Number of kernel launches this library has enqueued since load (all threads). */
long long libsmm_acc_b200_launch_count(void);
/* Run-time knobs of the FP64 stack kernels (dbcsr_b200/csrc/smm_tune.h; environment defaults DBCSR_B200_BALANCE / _ALIGN / _CHUNK /
 * _VARIANT): name = "balance" | "align" | "chunk" | "bigdmma" | "variant" | "trace_first" | "trace_count" | "seq"; get also answers "experiment"
 * (1 = library built with the kernel-variant table).  Return 0 / the value, -1 for an unknown name.
 * set_trace: device buffer of trace_count * 4096 * 128 64-bit words written by TRACE kernel variants (NULL = off). */
int libsmm_acc_b200_set_tunable(const char* name, long long value);
long long libsmm_acc_b200_get_tunable(const char* name);
void libsmm_acc_b200_set_trace(void* dev_words);
/* Programmatic-dependent-launch chain mode.  on != 0: the caller declares that `stream` (an acc stream handle) carries a chain of
 * independent stack drains -- between two libsmm_acc_process calls nothing that produces A, B, C or stack data is enqueued on it
 * except through full stream dependencies (memcpys, event waits).  The FP64 stack kernels then do not wait for their predecessor
 * grid before reading, so consecutive drains overlap tail and ramp-up; completion order stays stream order.  Default (and after
 * on == 0): every kernel waits for its predecessor before its first global read.  Returns 0, -2 for a NULL stream. */
int libsmm_acc_b200_stream_chain(void* stream, int on);
/* c_dbcsr_acc_memset_zero at a bounded rate: `nctas` CTAs (a handful) stream the zeros, so that a caller zeroing the NEXT multiply's C
 * buffer beside running stack kernels does not take the HBM bandwidth away from them in one burst.  offset, nbytes: multiples of 16. */
int libsmm_acc_b200_memset_zero_trickle(void* dev_mem, size_t offset, size_t nbytes, int nctas, void* stream);
/* Measured FP64 tensor-pipe (DMMA.8x8x4, register operands with random mantissas) throughput of the active device in GFLOP/s;
 * synchronises `stream`.
 * Introspection for roofline reports (bench.py); <= 0 on failure. */
double libsmm_acc_b200_fp64_peak_gflops(void* stream);
/* The same loop run back to back for `seconds` (0 < seconds <= 10): throughput over the second half -- the denominator for a kernel
 * timed inside a long step (power limits act within tens of milliseconds). */
double libsmm_acc_b200_fp64_peak_sustained_gflops(void* stream, double seconds);
/* Library identification string (static storage). */
const char* libsmm_acc_b200_version(void);

#if defined(__cplusplus)
}
#endif

#endif /* DBCSR_B200_ACC_LIBSMM_H */
