"""
dbcsr_b200/dbcsr.py -- host-side mirror of the reference's operator interface for the hot path: `dbcsr_multiply`
(src/mm/dbcsr_mm.F:336-1022, dbcsr_multiply_generic) for ONE process rank and real_8 data, with the reference's argument
meaning, limit handling and abort messages, on top of the accelerator engine (dbcsr_b200.multiply.DeviceMultiply -> C ABI).

    C := alpha * op(A) * op(B) + beta * C

What happens where (reference lines in brackets):
  * op(): block-wise transposes [dbcsr_mm.F:520-575, dbcsr_new_transposed]; symmetric / antisymmetric operands are expanded to
    full storage first [make_m2s -> dbcsr_desymmetrize_deep, dbcsr_mm_cannon.F:146-290];
  * limits first_row..last_k in FULL (element) indices, validation and "0 = no limit" normalisation [dbcsr_mm.F:629-700];
    the images are cropped to them: blocks outside are dropped, partially covered blocks keep zeros outside
    [make_images -> dbcsr_crop_matrix, ops/dbcsr_operations.F:1652-1833];
  * alpha is folded into the right image [dbcsr_mm.F:873-878], beta scales the existing C inside the limits [:706-709,
    dbcsr_scale with limits];
  * keep_product_data (beta != 0, retain_sparsity or row/column limits): the work matrices start from the existing C blocks
    [:701-705, :836-846], otherwise C is rebuilt from the product alone;
  * a product matrix with symmetry is computed in canonical (checkerboard) positions and skips the mirrored half
    [:711-715, dbcsr_make_index_canonical, dbcsr_mm_csr.F:280-292], then returns to upper-triangle storage;
  * filter_eps: on-the-fly norm filter + final filter of the product [dbcsr_mm_cannon.F:1038-1107, dbcsr_mm_multrec.F:700-758];
  * dbcsr_finalize: the per-thread work matrices are merged into one BCSR index (rows ascending, columns sorted)
    [work/dbcsr_work_operations.F:749-958].
The local multiply itself (stack building, kernels) is the engine's business; it is reached through a small backend object
so that the CPU test-suite can drive the identical pre/post-processing with the oracle standing in for the device.
There is NO CPU compute path in this module: without a backend the device engine is used and fails loudly without a GPU.
"""
import numpy as np

dbcsr_type_no_symmetry = "N"
dbcsr_type_symmetric = "S"
dbcsr_type_antisymmetric = "A"
dbcsr_no_transpose = "N"
dbcsr_transpose = "T"
dbcsr_conjugate_transpose = "C"


class DbcsrAbort(RuntimeError):
    """DBCSR_ABORT(message)"""


def checker_tr(row, column):
    """src/dist/dbcsr_dist_operations.F:65-75: must the block at (row, column) be stored transposed (checkerboard rule)?"""
    return bool((column + row) & 1) == (column >= row)


class DbcsrMatrix:
    """A finalized block-sparse matrix of one rank: BCSR index (row_p, col_i, blk_p; 1-based like DBCSR) over a flat real_8 data
    area with dense column-major blocks.  matrix_type 'S' / 'A' store the upper triangle only (row <= col)."""

    def __init__(self, name, row_blk_size, col_blk_size, matrix_type=dbcsr_type_no_symmetry):
        if matrix_type not in (dbcsr_type_no_symmetry, dbcsr_type_symmetric, dbcsr_type_antisymmetric):
            raise DbcsrAbort("Invalid matrix type")
        self.name = name
        self.row_blk_size = np.ascontiguousarray(row_blk_size, dtype=np.int32)
        self.col_blk_size = np.ascontiguousarray(col_blk_size, dtype=np.int32)
        self.matrix_type = matrix_type
        self.row_p = np.zeros(self.row_blk_size.size + 1, dtype=np.int64)
        self.col_i = np.zeros(0, dtype=np.int32)
        self.blk_p = np.zeros(0, dtype=np.int32)
        self.data = np.zeros(0)

    # ---- sizes
    @property
    def nblkrows_total(self):
        return int(self.row_blk_size.size)

    @property
    def nblkcols_total(self):
        return int(self.col_blk_size.size)

    @property
    def nfullrows_total(self):
        return int(self.row_blk_size.sum())

    @property
    def nfullcols_total(self):
        return int(self.col_blk_size.sum())

    @property
    def row_blk_offset(self):
        """1-based first full row of every block row, plus the end (dbcsr_row_block_offsets)."""
        return np.concatenate([[1], 1 + np.cumsum(self.row_blk_size, dtype=np.int64)])

    @property
    def col_blk_offset(self):
        return np.concatenate([[1], 1 + np.cumsum(self.col_blk_size, dtype=np.int64)])

    @property
    def nblks(self):
        return int(self.col_i.size)

    @property
    def nze(self):
        return int(self.data.size)

    def has_symmetry(self):
        return self.matrix_type != dbcsr_type_no_symmetry

    def get_occupation(self):
        """dbcsr_get_occupation: stored elements (mirrored off-diagonal blocks counted twice) / full size."""
        full = float(self.nfullrows_total) * float(self.nfullcols_total)
        if full == 0.0:
            return 0.0
        nze = float(self.nze)
        if self.has_symmetry():
            rows = self.block_rows()
            off = rows != self.col_i
            nze += float((self.row_blk_size[rows[off] - 1].astype(np.int64) * self.col_blk_size[self.col_i[off] - 1]).sum())
        return nze / full

    # ---- index helpers
    def block_rows(self):
        return np.repeat(np.arange(1, self.nblkrows_total + 1, dtype=np.int32), np.diff(self.row_p).astype(np.int64))

    def list3(self):
        """(row, col, blk_p) triples in BCSR order: the list index of an image (core/dbcsr_types.F:499-526)."""
        return np.ascontiguousarray(np.stack([self.block_rows(), self.col_i, self.blk_p], axis=1), dtype=np.int32)

    def block(self, i):
        """(m, n) view of stored block number i (0-based position in the index)."""
        r, c = int(self.block_rows()[i]), int(self.col_i[i])
        m, n = int(self.row_blk_size[r - 1]), int(self.col_blk_size[c - 1])
        o = int(self.blk_p[i]) - 1
        return self.data[o:o + m * n].reshape(n, m).T

    def blocks(self):
        """dict (row, col) -> (m, n) array (views into the data area)."""
        rows = self.block_rows()
        out = {}
        for i in range(self.nblks):
            r, c = int(rows[i]), int(self.col_i[i])
            m, n = int(self.row_blk_size[r - 1]), int(self.col_blk_size[c - 1])
            o = int(self.blk_p[i]) - 1
            out[(r, c)] = self.data[o:o + m * n].reshape(n, m).T
        return out

    @classmethod
    def from_blocks(cls, name, row_blk_size, col_blk_size, blocks, matrix_type=dbcsr_type_no_symmetry):
        """Build a finalized matrix from {(row, col): (m, n) array}; for 'S'/'A' only row <= col entries are accepted."""
        mat = cls(name, row_blk_size, col_blk_size, matrix_type)
        keys = sorted(blocks)
        counts = np.zeros(mat.nblkrows_total + 1, dtype=np.int64)
        col_i, blk_p, chunks, off = [], [], [], 1
        for (r, c) in keys:
            if not (1 <= r <= mat.nblkrows_total and 1 <= c <= mat.nblkcols_total):
                raise DbcsrAbort("Block coordinates out of range")
            if mat.has_symmetry() and r > c:
                raise DbcsrAbort("Symmetric matrices store the upper triangle")
            m, n = int(mat.row_blk_size[r - 1]), int(mat.col_blk_size[c - 1])
            b = np.asarray(blocks[(r, c)], dtype=np.float64)
            if b.shape != (m, n):
                raise DbcsrAbort("Block has wrong shape")
            counts[r] += 1
            col_i.append(c)
            blk_p.append(off)
            chunks.append(b.T.reshape(-1))  # column-major
            off += m * n
        mat.row_p = np.cumsum(counts)
        mat.col_i = np.array(col_i, dtype=np.int32)
        mat.blk_p = np.array(blk_p, dtype=np.int32)
        mat.data = np.concatenate(chunks) if chunks else np.zeros(0)
        return mat

    def copy(self, name=None):
        out = DbcsrMatrix(name or self.name, self.row_blk_size, self.col_blk_size, self.matrix_type)
        out.row_p, out.col_i, out.blk_p, out.data = self.row_p.copy(), self.col_i.copy(), self.blk_p.copy(), self.data.copy()
        return out

    def to_dense(self):
        """dbcsr_to_dense_local: full matrix, symmetry expanded (antisymmetric: mirrored blocks negated)."""
        ro, co = self.row_blk_offset - 1, self.col_blk_offset - 1
        dense = np.zeros((self.nfullrows_total, self.nfullcols_total))
        for (r, c), b in self.blocks().items():
            dense[ro[r - 1]:ro[r], co[c - 1]:co[c]] = b
            if self.has_symmetry() and r != c:
                dense[ro[c - 1]:ro[c], co[r - 1]:co[r]] = b.T if self.matrix_type == dbcsr_type_symmetric else -b.T
        return dense


# ------------------------------------------------------------------------------------------------ operand preparation
def _desymmetrized(mat):
    """dbcsr_desymmetrize_deep: every off-diagonal block of a symmetric / antisymmetric matrix also at its mirrored position."""
    if not mat.has_symmetry():
        return mat
    blocks = {}
    sign = 1.0 if mat.matrix_type == dbcsr_type_symmetric else -1.0
    for (r, c), b in mat.blocks().items():
        blocks[(r, c)] = b
        if r != c:
            blocks[(c, r)] = sign * b.T
    return DbcsrMatrix.from_blocks(mat.name, mat.row_blk_size, mat.col_blk_size, blocks)


def _transposed(mat):
    """dbcsr_new_transposed of a matrix without symmetry: block (r, c) -> block (c, r) transposed."""
    blocks = {(c, r): b.T for (r, c), b in mat.blocks().items()}
    return DbcsrMatrix.from_blocks(mat.name, mat.col_blk_size, mat.row_blk_size, blocks)


def _cropped_scaled(mat, row_bounds, col_bounds, scale=1.0):
    """dbcsr_crop_matrix (ops/dbcsr_operations.F:1652-1833) + scaling: blocks that intersect the full-index bounds
    (0 = unbounded) are kept, elements outside are zeroed; everything times `scale`."""
    f_row, l_row = row_bounds
    f_col, l_col = col_bounds
    if f_row == 0 and l_row == 0 and f_col == 0 and l_col == 0 and scale == 1.0:
        return mat
    ro, co = mat.row_blk_offset, mat.col_blk_offset
    f_row = f_row or 1
    l_row = l_row or mat.nfullrows_total
    f_col = f_col or 1
    l_col = l_col or mat.nfullcols_total
    blocks = {}
    for (r, c), b in mat.blocks().items():
        r0, r1 = int(ro[r - 1]), int(ro[r]) - 1  # full rows of the block, inclusive
        c0, c1 = int(co[c - 1]), int(co[c]) - 1
        if r1 < f_row or r0 > l_row or c1 < f_col or c0 > l_col:
            continue
        nb = np.zeros_like(b)
        i0, i1 = max(f_row, r0) - r0, min(l_row, r1) - r0 + 1
        j0, j1 = max(f_col, c0) - c0, min(l_col, c1) - c0 + 1
        nb[i0:i1, j0:j1] = scale * b[i0:i1, j0:j1]
        blocks[(r, c)] = nb
    return DbcsrMatrix.from_blocks(mat.name, mat.row_blk_size, mat.col_blk_size, blocks)


def _scale_within_limits(mat, beta, f_row, l_row, f_col, l_col):
    """dbcsr_scale(matrix, alpha_scalar, limits) (ops/dbcsr_operations.F): in place, only the elements inside the limits."""
    ro, co = mat.row_blk_offset, mat.col_blk_offset
    f_row = f_row or 1
    l_row = l_row or mat.nfullrows_total
    f_col = f_col or 1
    l_col = l_col or mat.nfullcols_total
    for (r, c), b in mat.blocks().items():
        r0, r1 = int(ro[r - 1]), int(ro[r]) - 1
        c0, c1 = int(co[c - 1]), int(co[c]) - 1
        if r1 < f_row or r0 > l_row or c1 < f_col or c0 > l_col:
            continue
        i0, i1 = max(f_row, r0) - r0, min(l_row, r1) - r0 + 1
        j0, j1 = max(f_col, c0) - c0, min(l_col, c1) - c0 + 1
        b[i0:i1, j0:j1] *= beta  # b is a view into mat.data


def dbcsr_finalize(row_blk_size, col_blk_size, parts, matrix_type=dbcsr_type_no_symmetry, name="product"):
    """dbcsr_finalize / dbcsr_merge_all (work/dbcsr_work_operations.F:749-958): merge the per-thread work matrices
    (rows, cols, blk_p, data) -- index in order of first touch -- into one BCSR matrix: rows ascending, columns sorted within a
    row, data area compacted in index order.  A block may appear in one work matrix only (threads own disjoint rows)."""
    out = DbcsrMatrix(name, row_blk_size, col_blk_size, matrix_type)
    rows = np.concatenate([np.asarray(p[0], dtype=np.int64) for p in parts]) if parts else np.zeros(0, dtype=np.int64)
    if rows.size == 0:
        return out
    cols = np.concatenate([np.asarray(p[1], dtype=np.int64) for p in parts])
    blk_p = np.concatenate([np.asarray(p[2], dtype=np.int64) for p in parts])
    part = np.concatenate([np.full(len(p[0]), i, dtype=np.int64) for i, p in enumerate(parts)])
    order = np.lexsort((cols, rows))
    rows, cols, blk_p, part = rows[order], cols[order], blk_p[order], part[order]
    if np.any((rows[1:] == rows[:-1]) & (cols[1:] == cols[:-1])):
        raise DbcsrAbort("Duplicate blocks in the work matrices")
    nze = out.row_blk_size[rows - 1].astype(np.int64) * out.col_blk_size[cols - 1]
    new_p = np.concatenate([[1], 1 + np.cumsum(nze)])
    # one vectorised gather: element e of sorted block i comes from (start of its part) + blk_p - 1 + e
    sizes = np.array([np.asarray(p[3]).size for p in parts], dtype=np.int64)
    base = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    data_all = np.concatenate([np.asarray(p[3], dtype=np.float64).reshape(-1) for p in parts])
    src0 = base[part] + blk_p - 1
    if np.any(src0 + nze > base[part] + sizes[part]):
        raise DbcsrAbort("Work matrix index points outside its data area")
    total = int(new_p[-1] - 1)
    if np.array_equal(src0, new_p[:-1] - 1) and data_all.size == total:
        data = data_all  # already in BCSR order and compact (finalized on the device)
    else:
        data = data_all[np.repeat(src0 - (new_p[:-1] - 1), nze) + np.arange(total, dtype=np.int64)]
    out.row_p = np.concatenate([[0], np.cumsum(np.bincount(rows - 1, minlength=out.nblkrows_total))])
    out.col_i = cols.astype(np.int32)
    out.blk_p = new_p[:-1].astype(np.int32)
    out.data = data
    return out


def dbcsr_checksum(matrix, pos=False):
    """dbcsr_checksum (src/dist/dbcsr_dist_util.F:432-547, pd_blk_cs :549-575) of a finalized real_8 matrix: the sum of the squared
    elements, or - position dependent - the sum of x(r,c) * log|r * c| over the full (1-based) element coordinates.  Summed block by
    block and row by row like the reference (the order only matters at the 1e-16 level)."""
    ro, co = matrix.row_blk_offset, matrix.col_blk_offset
    rows = matrix.block_rows()
    total, row_sum, cur_row = 0.0, 0.0, -1
    for i in range(matrix.nblks):
        r, c = int(rows[i]), int(matrix.col_i[i])
        if r != cur_row:
            total += row_sum
            row_sum, cur_row = 0.0, r
        b = matrix.block(i)
        if pos:
            rr = np.arange(int(ro[r - 1]), int(ro[r]), dtype=np.float64)[:, None]
            cc = np.arange(int(co[c - 1]), int(co[c]), dtype=np.float64)[None, :]
            row_sum += float((b * np.log(np.abs(rr * cc))).sum())
        else:
            row_sum += float((b * b).sum())
    return total + row_sum


def dbcsr_scale(matrix, alpha_scalar, limits=None):
    """dbcsr_scale (ops/dbcsr_operations.F): in place; limits = (first_row, last_row, first_col, last_col) in full indices, 0 = open."""
    f_row, l_row, f_col, l_col = limits if limits is not None else (0, 0, 0, 0)
    _scale_within_limits(matrix, float(alpha_scalar), f_row, l_row, f_col, l_col)


# ------------------------------------------------------------------------------------------------ backends
class DeviceBackend:
    """The product path: panels go to the device, the host engine builds the stacks, libsmm_acc_process drains them
    (dbcsr_b200.multiply.DeviceMultiply).  Needs the built C-ABI library and a GPU; there is no fallback."""

    def __init__(self, acc, nthreads=1, cfg=None, device_build=False):
        """device_build: stacks and C index are built on the device (include/dbcsr_b200_host.h, DBCSR_B200_DEVICE_BUILD) instead of by
        the host threads; results are identical."""
        self.acc, self.nthreads, self.cfg, self.device_build = acc, nthreads, cfg, device_build

    def local_multiply(self, m_sizes, n_sizes, k_sizes, left, right, c_preset, keep_sparsity, c_symmetry, filter_eps, final_filter):
        from . import host
        from .multiply import DeviceMultiply

        dm = DeviceMultiply(self.acc, m_sizes, n_sizes, k_sizes, left.data.size, right.data.size, right.nblks,
                            nthreads=self.nthreads, cfg=self.cfg, mode=host.LAUNCH | (host.DEVICE_BUILD if self.device_build else 0))
        try:
            b_list = right.list3()
            dm.upload_panels(np.ascontiguousarray(left.data), np.ascontiguousarray(right.data), b_list)
            dm.multiply(left.list3(), b_list, filter_eps=filter_eps, c_preset=c_preset, retain_sparsity=keep_sparsity,
                        c_symmetry=c_symmetry)
            # dbcsr_finalize on the device: final filter (if any), BCSR order, compaction; the download is the final data area
            dm.finalize_c(filter_eps if final_filter else None)
            prod = dm.download_c()
            return [(r, c, p, np.array(d, copy=True)) for (r, c, p, d) in prod.parts], dm.engine.flop()
        finally:
            dm.close()


_default_backend = None


def set_default_backend(backend):
    """Backend used by dbcsr_multiply when none is passed (e.g. DeviceBackend(Acc(0), nthreads=8))."""
    global _default_backend
    _default_backend = backend


# ------------------------------------------------------------------------------------------------ the operator
def dbcsr_multiply(transa, transb, alpha, matrix_a, matrix_b, beta, matrix_c, first_row=None, last_row=None, first_column=None,
                   last_column=None, first_k=None, last_k=None, retain_sparsity=None, filter_eps=None, backend=None):
    """C := alpha * op(A) * op(B) + beta * C  (src/mm/dbcsr_mm.F:336-1022).  matrix_c is updated in place; returns flop
    (2*m*n*k summed over the block products actually performed).  Arguments as in the reference; limits are 1-based FULL
    indices.  Aborts (DbcsrAbort) carry the reference's messages."""
    global _default_backend
    if matrix_a.get_occupation() > 1:
        raise DbcsrAbort("Matrix A occupation > 1")
    if matrix_b.get_occupation() > 1:
        raise DbcsrAbort("Matrix B occupation > 1")
    if matrix_c.get_occupation() > 1:
        raise DbcsrAbort("Matrix C occupation > 1")
    transa_l, transb_l = str(transa).upper(), str(transb).upper()

    # ---- op(A), op(B) (real data: 'C' = 'T')
    if transa_l == dbcsr_no_transpose:
        matrix_left = _desymmetrized(matrix_a)
    elif transa_l in (dbcsr_transpose, dbcsr_conjugate_transpose):
        matrix_left = _transposed(_desymmetrized(matrix_a))
    else:
        raise DbcsrAbort("wrong transa_l = " + transa_l)
    if transb_l == dbcsr_no_transpose:
        matrix_right = _desymmetrized(matrix_b)
    elif transb_l in (dbcsr_transpose, dbcsr_conjugate_transpose):
        matrix_right = _transposed(_desymmetrized(matrix_b))
    else:
        raise DbcsrAbort("wrong transb_l = " + transb_l)
    if not np.array_equal(matrix_c.row_blk_offset, matrix_left.row_blk_offset):
        raise DbcsrAbort("C/A rows not equal")
    if not np.array_equal(matrix_c.col_blk_offset, matrix_right.col_blk_offset):
        raise DbcsrAbort("C/B columns not equal")
    if not np.array_equal(matrix_left.col_blk_offset, matrix_right.row_blk_offset):
        raise DbcsrAbort("A cols/B rows not equal")

    # ---- limits (dbcsr_mm.F:629-700)
    f_row, l_row = 1, matrix_c.nfullrows_total
    f_col, l_col = 1, matrix_c.nfullcols_total
    f_k, l_k = 1, matrix_left.nfullcols_total
    if first_row is not None:
        if first_row < 1 or first_row > matrix_c.nfullrows_total:
            raise DbcsrAbort("Invalid first row specified")
        f_row = first_row
    if last_row is not None:
        if last_row > matrix_c.nfullrows_total:
            raise DbcsrAbort("Invalid last row specified")
        l_row = last_row
    if first_column is not None:
        if first_column < 1 or first_column > matrix_c.nfullcols_total:
            raise DbcsrAbort("Invalid first col specified")
        f_col = first_column
    if last_column is not None:
        if last_column > matrix_c.nfullcols_total:
            raise DbcsrAbort("Invalid last column specified (C)")
        if last_column > matrix_right.nfullcols_total:
            raise DbcsrAbort("Invalid last column specified (B)")
        l_col = last_column
    if first_k is not None:
        if first_k < 1 or first_k > matrix_left.nfullcols_total:
            raise DbcsrAbort("Invalid first k specified (A)")
        f_k = first_k
    if last_k is not None:
        if last_k > matrix_left.nfullcols_total:
            raise DbcsrAbort("Invalid last k specified (A)")
        l_k = last_k
    # 0 = no limit
    if f_row == 1:
        f_row = 0
    if l_row == matrix_left.nfullrows_total:
        l_row = 0
    if f_col == 1:
        f_col = 0
    l_col = min(l_col, matrix_right.nfullcols_total, matrix_c.nfullcols_total)
    if f_col <= 1 and l_col == matrix_right.nfullcols_total and matrix_right.nfullcols_total == matrix_c.nfullcols_total:
        l_col = 0
    if f_k == 1:
        f_k = 0
    if l_k == matrix_left.nfullcols_total:
        l_k = 0
    if f_row > l_row and l_row > 0:
        raise DbcsrAbort("Last row smaller than first row")
    if f_col > l_col and l_col > 0:
        raise DbcsrAbort("Last col smaller than first col")

    # ---- product data kept?  beta scaling (dbcsr_mm.F:701-709)
    keep_sparsity = bool(retain_sparsity) if retain_sparsity is not None else False
    keep_product_data = (keep_sparsity or beta != 0.0 or (0 < l_col < matrix_c.nfullcols_total)
                         or (0 < l_row < matrix_c.nfullrows_total))
    if beta != 1.0 and keep_product_data:
        _scale_within_limits(matrix_c, beta, f_row, l_row, f_col, l_col)
    product_reindex = matrix_c.has_symmetry()

    # ---- images: crop to the limits, alpha into the right one (make_m2s / make_images)
    left = _cropped_scaled(matrix_left, (f_row, l_row), (f_k, l_k))
    right = _cropped_scaled(matrix_right, (f_k, l_k), (f_col, l_col), scale=float(alpha))

    # ---- existing product blocks -> work matrix (canonical positions when C has symmetry)
    c_preset = None
    if keep_product_data:
        rows, cols, chunks = [], [], []
        sign = -1.0 if matrix_c.matrix_type == dbcsr_type_antisymmetric else 1.0
        for (r, c), b in sorted(_canonical_blocks(matrix_c, product_reindex, sign).items()):
            rows.append(r)
            cols.append(c)
            chunks.append(np.ascontiguousarray(b.T).reshape(-1))
        c_preset = (np.array(rows, dtype=np.int32), np.array(cols, dtype=np.int32), np.concatenate(chunks) if chunks else np.zeros(0))

    # ---- local multiply on the accelerator
    if backend is None:
        backend = _default_backend
    if backend is None:
        from . import lib as acclib

        backend = _default_backend = DeviceBackend(acclib.Acc(0))
    use_filter = filter_eps is not None
    if left.nblks == 0 or right.nblks == 0:
        parts, flop = ([(c_preset[0], c_preset[1], _running_offsets(matrix_c, c_preset), c_preset[2])] if c_preset is not None else []), 0
    else:
        parts, flop = backend.local_multiply(matrix_c.row_blk_size, matrix_c.col_blk_size, matrix_left.col_blk_size, left, right,
                                             c_preset, keep_sparsity, product_reindex, filter_eps if use_filter else None,
                                             use_filter and not keep_sparsity)

    # ---- dbcsr_finalize, back to the stored form of C (dbcsr_mm.F:925-985)
    product = dbcsr_finalize(matrix_c.row_blk_size, matrix_c.col_blk_size, parts)
    if product_reindex:
        sign = -1.0 if matrix_c.matrix_type == dbcsr_type_antisymmetric else 1.0
        blocks = {}
        for (r, c), b in product.blocks().items():
            if r > c:
                if (c, r) in blocks:
                    raise DbcsrAbort("Both halves of a symmetric product block were computed")
                blocks[(c, r)] = sign * b.T
            else:
                if (r, c) in blocks:
                    raise DbcsrAbort("Both halves of a symmetric product block were computed")
                blocks[(r, c)] = b
        product = DbcsrMatrix.from_blocks(matrix_c.name, matrix_c.row_blk_size, matrix_c.col_blk_size, blocks, matrix_c.matrix_type)
    matrix_c.row_p, matrix_c.col_i, matrix_c.blk_p, matrix_c.data = product.row_p, product.col_i, product.blk_p, product.data
    if matrix_c.nblks > matrix_c.nblkrows_total * matrix_c.nblkcols_total:
        raise DbcsrAbort("Bug: Matrix contains too many blocks")
    return int(flop)


def _canonical_blocks(matrix_c, reindex, sign):
    """dbcsr_make_index_canonical: blocks of a matrix with symmetry move to the position the checkerboard rule computes
    (transposed, and negated for antisymmetric matrices, when that is the mirrored position)."""
    if not reindex:
        return {k: v for k, v in matrix_c.blocks().items()}
    out = {}
    for (r, c), b in matrix_c.blocks().items():
        if r != c and checker_tr(r, c):
            out[(c, r)] = sign * b.T
        else:
            out[(r, c)] = b
    return out


def _running_offsets(matrix_c, c_preset):
    rows, cols, _ = c_preset
    nze = matrix_c.row_blk_size[rows - 1].astype(np.int64) * matrix_c.col_blk_size[cols - 1]
    return (1 + np.concatenate([[0], np.cumsum(nze)[:-1]])).astype(np.int32) if rows.size else np.zeros(0, dtype=np.int32)
