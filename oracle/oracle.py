"""
oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end to oracle/liboracle.so (our C restatement, oracle/dbcsr_oracle.c) and, when present, to
oracle/_ref/libref_smm.so (the reference's own CPU checker functions compiled from /root/reference).
Only tests/, bench.py (cpu_baseline / --impl reference) and __graft_entry__.smoke() import this module.
"""
import ctypes
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_c_int, _c_long, _c_double, _c_voidp = ctypes.c_int, ctypes.c_long, ctypes.c_double, ctypes.c_void_p


def build(force=False):
    """Compile oracle/liboracle.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "dbcsr_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/src/acc/libsmm_acc"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return so


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        so = build()
        L = ctypes.CDLL(so)
        L.orc_stack_calc.argtypes = [_i32p, _c_int, _f64p, _f64p, _f64p, _c_int, _c_int, _c_int]
        L.orc_host_stack.argtypes = [_i32p, _c_int, _f64p, _f64p, _f64p, _c_voidp]
        L.orc_host_stacks_threaded.argtypes = [_i32p, _i64p, _i32p, _c_int, _c_int, _f64p, _f64p, _f64p, _c_voidp]
        L.orc_max_threads.restype = _c_int
        L.orc_transpose.argtypes = [_i32p, _c_int, _f64p, _c_int, _c_int]
        L.orc_norms.argtypes = [_f64p, _c_int, _i32p, _i32p, _f32p]
        L.orc_mat_init.argtypes = [_f64p, _c_int, _c_int, _c_int, _c_int]
        L.orc_stack_init.argtypes = [_i32p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int]
        L.orc_checksum.argtypes = [_f64p, _c_int, _c_int, _c_int]
        L.orc_checksum.restype = _c_double
        L.orc_checksum_transp.argtypes = [_f64p, _c_int, _c_int, _c_int]
        L.orc_checksum_transp.restype = _c_double
        L.orc_set_larnv_seed.argtypes = [_c_int, _c_int, _c_int, _c_int, _c_int, _i32p]
        L.orc_dlarnv1.argtypes = [_i32p, _c_int, _f64p]
        L.orc_random_blocks.argtypes = [_c_int, _c_int, _c_double, _c_int, _i32p, _i32p, _c_long]
        L.orc_random_blocks.restype = _c_long
        L.orc_fill_blocks.argtypes = [_c_long, _i32p, _i32p, _i64p, _i32p, _i32p, _c_int, _c_int, _c_int, _f64p]
        L.orc_dbcsr_checksum.argtypes = [_c_long, _i32p, _i32p, _i64p, _i32p, _i32p, _i32p, _i32p, _f64p, _c_int]
        L.orc_dbcsr_checksum.restype = _c_double
        L.orc_block_gemm.argtypes = [_c_int, _c_int, _c_int, _f64p, _c_int, _f64p, _f64p]
        _lib = L
    return _lib


def ref():
    """The reference's own checker functions (None when oracle/_ref was not built/shipped)."""
    global _ref
    if _ref is None:
        so = os.path.join(_HERE, "_ref", "libref_smm.so")
        if not os.path.exists(so):
            return None
        R = ctypes.CDLL(so)
        R.ref_matInit.argtypes = [_f64p, _c_int, _c_int, _c_int, _c_int]
        R.ref_stackInit.argtypes = [_i32p] + [_c_int] * 7
        R.ref_stackCalc.argtypes = [_i32p, _c_int, _f64p, _f64p, _f64p, _c_int, _c_int, _c_int]
        R.ref_stackTransp.argtypes = [_i32p, _c_int, _f64p, _f64p, _c_int, _c_int]
        R.ref_checkSum.argtypes = [_f64p, _c_int, _c_int, _c_int]
        R.ref_checkSum.restype = _c_double
        R.ref_checkSumTransp.argtypes = [_f64p, _c_int, _c_int, _c_int]
        R.ref_checkSumTransp.restype = _c_double
        _ref = R
    return _ref


_libc = ctypes.CDLL(None)


def srand(seed):
    _libc.srand(ctypes.c_uint(seed))


# ------------------------------------------------------------------ BLAS for the CPU baseline
_blas = None


def openblas():
    """scipy's bundled OpenBLAS (third-party BLAS = what the reference's DGEMM resolves to)."""
    global _blas
    if _blas is None:
        import scipy

        cands = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so"))
        if not cands:
            return None
        _blas = ctypes.CDLL(cands[0])
        try:
            _blas.scipy_openblas_set_num_threads.argtypes = [_c_int]
            _blas.scipy_openblas_set_num_threads(1)  # DBCSR threads at the OpenMP level, BLAS stays serial
        except AttributeError:
            pass
    return _blas


def dgemm_ptr():
    b = openblas()
    if b is None:
        return None
    return ctypes.cast(b.scipy_dgemm_, _c_voidp)


def lapack_dlarnv(iseed, n):
    """Reference LAPACK dlarnv(idist=1) from scipy's OpenBLAS: pins orc_dlarnv1."""
    b = openblas()
    f = b.scipy_dlarnv_
    idist = ctypes.c_int(1)
    nn = ctypes.c_int(n)
    seed = np.array(iseed, dtype=np.int32)
    x = np.empty(n, dtype=np.float64)
    f(ctypes.byref(idist), seed.ctypes.data_as(_c_voidp), ctypes.byref(nn), x.ctypes.data_as(_c_voidp))
    return x, seed


# ------------------------------------------------------------------ numpy-level helpers
def stack_calc(stack3, c, a, b, m, n, k):
    stack3 = np.ascontiguousarray(stack3, dtype=np.int32).reshape(-1)
    lib().orc_stack_calc(stack3, stack3.size // 3, c, a, b, m, n, k)
    return c


def host_stack(params7, a, b, c, use_blas=False):
    params7 = np.ascontiguousarray(params7, dtype=np.int32).reshape(-1)
    lib().orc_host_stack(params7, params7.size // 7, a, b, c, dgemm_ptr() if use_blas else None)
    return c


def transpose_blocks(stack, mat, m, n):
    stack = np.ascontiguousarray(stack, dtype=np.int32)
    lib().orc_transpose(stack, stack.size, mat, m, n)
    return mat


def norms(mat, offsets, nelems):
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    nelems = np.ascontiguousarray(nelems, dtype=np.int32)
    out = np.empty(offsets.size, dtype=np.float32)
    lib().orc_norms(mat, offsets.size, offsets, nelems, out)
    return out


class BlockMatrix:
    """Block-sparse matrix in the layout DBCSR hands to the multiply: BCSR-ordered block list + flat data area.

    rows/cols: 1-based block coordinates (BCSR order); offsets: 0-based element offset of each block (col-major
    blocks, contiguous); row_blk_size / col_blk_size per block row / col.
    """

    def __init__(self, row_blk_size, col_blk_size, rows, cols, data=None):
        self.row_blk_size = np.ascontiguousarray(row_blk_size, dtype=np.int32)
        self.col_blk_size = np.ascontiguousarray(col_blk_size, dtype=np.int32)
        self.rows = np.ascontiguousarray(rows, dtype=np.int32)
        self.cols = np.ascontiguousarray(cols, dtype=np.int32)
        nze = self.row_blk_size[self.rows - 1].astype(np.int64) * self.col_blk_size[self.cols - 1].astype(np.int64)
        self.offsets = np.zeros(self.rows.size, dtype=np.int64)
        if self.rows.size:
            self.offsets[1:] = np.cumsum(nze)[:-1]
        self.nze = int(nze.sum())
        self.data = np.zeros(self.nze, dtype=np.float64) if data is None else data
        self.row_off = np.concatenate([[1], 1 + np.cumsum(self.row_blk_size)[:-1]]).astype(np.int32)
        self.col_off = np.concatenate([[1], 1 + np.cumsum(self.col_blk_size)[:-1]]).astype(np.int32)

    @property
    def nblks(self):
        return int(self.rows.size)

    def index_list(self):
        """(row, col, blk_p) with 1-based element offsets = the coo_l list index of an untransposed panel."""
        return [(int(r), int(c), int(o) + 1) for r, c, o in zip(self.rows, self.cols, self.offsets)]

    def checksum(self, pos=False):
        return lib().orc_dbcsr_checksum(self.nblks, self.rows, self.cols, self.offsets, self.row_blk_size, self.col_blk_size,
                                        self.row_off, self.col_off, self.data, 1 if pos else 0)

    def to_dense(self):
        M, N = int(self.row_blk_size.sum()), int(self.col_blk_size.sum())
        d = np.zeros((M, N))
        for r, c, o in zip(self.rows, self.cols, self.offsets):
            m, n = self.row_blk_size[r - 1], self.col_blk_size[c - 1]
            d[self.row_off[r - 1] - 1:self.row_off[r - 1] - 1 + m, self.col_off[c - 1] - 1:self.col_off[c - 1] - 1 + n] = \
                self.data[o:o + m * n].reshape(n, m).T
        return d


def random_matrix(row_blk_size, col_blk_size, sparsity, counter):
    """dbcsr_make_random_matrix (src/ops/dbcsr_test_methods.F:318-465), symmetry 'N', one rank.
    counter = value of randmat_counter for this matrix (12341313 + number of matrices made since the reset)."""
    nrow, ncol = len(row_blk_size), len(col_blk_size)
    cap = int(nrow * ncol * (1.0 - (sparsity / 100.0 if sparsity > 1 else sparsity)) * 1.2) + 1024
    while True:
        rows = np.empty(cap, dtype=np.int32)
        cols = np.empty(cap, dtype=np.int32)
        cnt = lib().orc_random_blocks(nrow, ncol, float(sparsity), counter, rows, cols, cap)
        if cnt >= 0:
            break
        cap = -cnt - 1
    mat = BlockMatrix(row_blk_size, col_blk_size, rows[:cnt].copy(), cols[:cnt].copy())
    lib().orc_fill_blocks(mat.nblks, mat.rows, mat.cols, mat.offsets, mat.row_blk_size, mat.col_blk_size, nrow, ncol, counter,
                          mat.data)
    return mat


def random_block_sizes(size_sum, size_mix):
    """dbcsr_make_random_block_sizes, src/ops/dbcsr_test_methods.F:466-510. size_mix = [mult1, size1, mult2, size2, ...]."""
    nmix = len(size_mix) // 2
    mult = size_mix[0::2]
    sz = size_mix[1::2]
    cnt = [1] * nmix
    out, cur, sel = [], 0, 0
    while cur < size_sum:
        bs = min(sz[sel], size_sum - cur)
        out.append(bs)
        cur += bs
        cnt[sel] += 1
        if cnt[sel] > mult[sel]:
            cnt[sel] = 1
            sel = (sel + 1) % nmix
    return out


def multiply_blocks(A, B, C_in=None, transa=False):
    """Reference result of C_out = op(A)*B + C_in on the block level (plain loops via orc_block_gemm).
    Used only by the golden-checksum test and as the dense-free checker for small cases."""
    L = lib()
    if transa:
        m_sizes, k_sizes = A.col_blk_size, A.row_blk_size
    else:
        m_sizes, k_sizes = A.row_blk_size, A.col_blk_size
    n_sizes = B.col_blk_size
    blocks = {}
    if C_in is not None:
        for r, c, o in zip(C_in.rows, C_in.cols, C_in.offsets):
            nz = int(m_sizes[r - 1]) * int(n_sizes[c - 1])
            blocks[(int(r), int(c))] = C_in.data[o:o + nz].copy()
    b_by_row = {}
    for idx, (r, c) in enumerate(zip(B.rows, B.cols)):
        b_by_row.setdefault(int(r), []).append(idx)
    for ia, (ar, ac) in enumerate(zip(A.rows, A.cols)):
        i, kk = (int(ac), int(ar)) if transa else (int(ar), int(ac))
        m, k = int(m_sizes[i - 1]), int(k_sizes[kk - 1])
        ablk = A.data[A.offsets[ia]:A.offsets[ia] + m * k]
        for ib in b_by_row.get(kk, []):
            j = int(B.cols[ib])
            n = int(n_sizes[j - 1])
            cb = blocks.get((i, j))
            if cb is None:
                cb = np.zeros(m * n)
                blocks[(i, j)] = cb
            L.orc_block_gemm(m, n, k, ablk, 1 if transa else 0, B.data[B.offsets[ib]:B.offsets[ib] + k * n], cb)
    keys = sorted(blocks.keys())
    C = BlockMatrix(m_sizes, n_sizes, [k_[0] for k_ in keys], [k_[1] for k_ in keys])
    for (key, o) in zip(keys, C.offsets):
        C.data[o:o + blocks[key].size] = blocks[key]
    return C
