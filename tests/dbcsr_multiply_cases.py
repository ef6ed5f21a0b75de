"""Shared driver of the dbcsr_multiply tests (CPU: oracle backend, GPU: device backend), modelled on the reference's
tests/dbcsr_test_multiply.F: dbcsr_test_multiplies (:66-346: symmetry table, transposes, B = +-A for products with symmetry),
test_multiply (:348-521) and dbcsr_check_multiply (:523-788: dense GEMM on the limited sub-matrices, criterion
||C_dbcsr - C_dense||_oo / ((||A||_oo + ||B||_oo + ||C_in||_oo) * n * eps) <= 10); cases from tests/dbcsr_unittest1.F:95-330, dbcsr_unittest2.F:80-107, dbcsr_unittest3.F:79-120."""
import numpy as np

from dbcsr_b200 import dbcsr as D
from oracle import oracle as orc

# (a_symm, b_symm, c_symm): tests/dbcsr_test_multiply.F:98-121
SYMMETRIES = [("N", "N", "N"), ("S", "N", "N"), ("A", "N", "N"), ("N", "S", "N"), ("S", "S", "N"), ("A", "S", "N"), ("N", "A", "N"),
              ("S", "A", "N"), ("A", "A", "N"), ("N", "N", "S"), ("S", "S", "S"), ("A", "A", "S")]

# name, matrix_sizes, sparsities, retain_sparsity, alpha, beta, bs_m, bs_n, bs_k, limits  (real part of the reference's scalars)
UNITTEST1_CASES = [
    ("multiply_ALPHA", (20, 20, 20), (0.5, 0.5, 0.5), True, -3.0, 0.0, [1, 4], [1, 4], [1, 4], (2, 6, 3, 7, 6, 7)),
    ("multiply_BETA", (20, 20, 20), (0.5, 0.5, 0.5), True, 1.0, 3.0, [1, 4], [1, 4], [1, 4], (2, 6, 3, 7, 6, 7)),
    ("multiply_LIMITS_COL_1", (50, 50, 50), (0.0, 0.0, 0.0), False, 1.0, 0.0, [1, 2], [1, 2], [1, 2], (1, 50, 1, 20, 1, 50)),
    ("multiply_LIMITS_COL_2", (50, 50, 50), (0.0, 0.0, 0.0), False, 1.0, 0.0, [1, 2], [1, 2], [1, 2], (1, 50, 9, 18, 1, 50)),
    ("multiply_LIMITS_COL_3", (50, 50, 50), (0.5, 0.5, 0.5), True, 1.0, 0.0, [1, 2], [1, 2], [1, 2], (1, 50, 9, 18, 1, 50)),
    ("multiply_LIMITS_COL_4", (25, 50, 75), (0.5, 0.5, 0.5), True, 1.0, 0.0, [1, 2], [1, 2], [1, 2], (1, 25, 9, 18, 1, 75)),
    ("multiply_LIMITS_K_1", (50, 50, 50), (0.0, 0.0, 0.0), False, 1.0, 0.0, [1, 2], [1, 2], [1, 2], (1, 50, 1, 50, 1, 20)),
    ("multiply_LIMITS_K_2", (50, 50, 50), (0.0, 0.0, 0.0), False, 1.0, 0.0, [1, 2], [1, 2], [1, 2], (1, 50, 1, 50, 9, 18)),
    ("multiply_LIMITS_K_3", (50, 50, 50), (0.5, 0.5, 0.5), True, 1.0, 0.0, [1, 2], [1, 2], [1, 2], (1, 50, 1, 50, 9, 18)),
    ("multiply_LIMITS_K_4", (25, 50, 75), (0.5, 0.5, 0.5), True, 1.0, 0.0, [1, 2], [1, 2], [1, 2], (1, 25, 1, 50, 9, 18)),
    ("multiply_LIMITS_MIX_1", (50, 50, 50), (0.0, 0.0, 0.0), False, 1.0, 0.0, [1, 2], [1, 2], [1, 2], (9, 18, 11, 20, 1, 50)),
    ("multiply_LIMITS_MIX_2", (50, 50, 50), (0.0, 0.0, 0.0), False, 1.0, 0.0, [1, 2], [1, 2], [1, 2], (1, 50, 9, 10, 11, 20)),
    ("multiply_LIMITS_MIX_3", (50, 50, 50), (0.0, 0.0, 0.0), False, 1.0, 0.0, [1, 2], [1, 2], [1, 2], (9, 20, 1, 50, 11, 18)),
    ("multiply_LIMITS_MIX_4", (50, 50, 50), (0.0, 0.0, 0.0), False, 1.0, 0.0, [1, 2], [1, 2], [1, 2], (11, 20, 11, 20, 13, 18)),
    ("multiply_LIMITS_MIX_5", (50, 50, 50), (0.5, 0.5, 0.5), True, 1.0, 0.0, [1, 2], [1, 2], [1, 2], (11, 20, 11, 20, 13, 18)),
    ("multiply_LIMITS_MIX_6", (25, 50, 75), (0.5, 0.5, 0.5), True, 1.0, 0.0, [1, 2], [1, 2], [1, 2], (11, 20, 11, 20, 13, 18)),
    ("multiply_LIMITS_MIX_7", (25, 50, 75), (0.5, 0.5, 0.5), True, 1.0, 0.0, [1, 2], [1, 2, 1, 3], [1, 3, 1, 2, 1, 0], (11, 20, 11, 20, 6, 10)),
    ("multiply_LIMITS_ROW_1", (50, 50, 50), (0.0, 0.0, 0.0), False, 1.0, 0.0, [1, 2], [1, 2], [1, 2], (1, 20, 1, 50, 1, 50)),
    ("multiply_LIMITS_ROW_2", (50, 50, 50), (0.0, 0.0, 0.0), False, 1.0, 0.0, [1, 2], [1, 2], [1, 2], (9, 18, 1, 50, 1, 50)),
    ("multiply_LIMITS_ROW_3", (50, 50, 50), (0.5, 0.5, 0.5), True, 1.0, 0.0, [1, 2], [1, 2], [1, 2], (9, 18, 1, 50, 1, 50)),
    # tests/dbcsr_unittest2.F:80-107 (large blocks, rectangular matrices) and tests/dbcsr_unittest3.F:79-120 (the GPU-targeted block mixes)
    ("large_blocks_1", (500, 500, 500), (0.5, 0.5, 0.5), False, 1.0, 0.0, [1, 100], [1, 100], [1, 100], (1, 500, 1, 500, 1, 500)),
    ("large_blocks_2", (500, 50, 50), (0.5, 0.5, 0.5), False, 1.0, 0.0, [1, 100], [1, 10], [1, 10], (1, 500, 1, 50, 1, 50)),
    ("rectangular_matrix_M", (500, 50, 50), (0.5, 0.5, 0.5), False, 1.0, 0.0, [1, 5], [1, 5], [1, 5], (1, 500, 1, 50, 1, 50)),
    ("rectangular_matrix_K", (50, 50, 500), (0.5, 0.5, 0.5), False, 1.0, 0.0, [1, 5], [1, 5], [1, 5], (1, 50, 1, 50, 1, 500)),
    ("blocks_1_3_4", (496, 48, 48), (0.5, 0.5, 0.5), False, 1.0, 0.0, [1, 1, 1, 3, 1, 4], [1, 1, 1, 3, 1, 4], [1, 1, 1, 3, 1, 4], (1, 496, 1, 48, 1, 48)),
    ("blocks_4_5_7", (496, 48, 48), (0.5, 0.5, 0.5), False, 1.0, 0.0, [1, 4, 1, 5, 1, 7], [1, 4, 1, 5, 1, 7], [1, 4, 1, 5, 1, 7], (1, 496, 1, 48, 1, 48)),
    ("blocks_5_8_9", (506, 44, 44), (0.5, 0.5, 0.5), False, 1.0, 0.0, [1, 5, 1, 8, 1, 9], [1, 5, 1, 8, 1, 9], [1, 5, 1, 8, 1, 9], (1, 506, 1, 44, 1, 44)),
    ("blocks_4_13_25", (504, 42, 42), (0.5, 0.5, 0.5), False, 1.0, 0.0, [1, 4, 1, 13, 1, 25], [1, 4, 1, 13, 1, 25], [1, 4, 1, 13, 1, 25],
     (1, 504, 1, 42, 1, 42)),
    ("blocks_14_29_32", (525, 75, 75), (0.5, 0.5, 0.5), False, 1.0, 0.0, [1, 14, 1, 29, 1, 32], [1, 14, 1, 29, 1, 32], [1, 14, 1, 29, 1, 32],
     (1, 525, 1, 75, 1, 75)),
    ("blocks_H2O", (552, 46, 46), (0.5, 0.5, 0.5), False, 1.0, 0.0, [1, 23], [1, 23], [1, 23], (1, 552, 1, 46, 1, 46)),
    ("blocks_45_67_78", (570, 190, 190), (0.5, 0.5, 0.5), False, 1.0, 0.0, [1, 45, 1, 67, 1, 78], [1, 45, 1, 67, 1, 78], [1, 45, 1, 67, 1, 78],
     (1, 570, 1, 190, 1, 190)),
    # full-range multiplies on the hot path's block sizes (tests/dbcsr_unittest3.F:76-118 mixes), also exercising symmetric products
    ("multiply_SQUARE_23", (115, 115, 115), (0.4, 0.4, 0.6), False, 1.0, 0.0, [1, 23], [1, 23], [1, 23], (1, 115, 1, 115, 1, 115)),
    ("multiply_MIX_5_13_23", (87, 87, 87), (0.5, 0.5, 0.5), False, 2.0, 1.0, [1, 5, 1, 13, 1, 23], [1, 5, 1, 13, 1, 23], [1, 5, 1, 13, 1, 23],
     (1, 87, 1, 87, 1, 87)),
    ("multiply_RETAIN_MIX", (87, 64, 70), (0.5, 0.5, 0.3), True, 1.0, 2.0, [1, 5, 1, 13], [1, 13, 1, 5], [1, 23, 1, 5], (1, 87, 1, 64, 1, 70)),
]


def random_matrix(name, row_sizes, col_sizes, sparsity, symmetry, rng):
    """Random block pattern with the given sparsity (fraction of ABSENT blocks) and uniform(-1,1) values; symmetric /
    antisymmetric matrices get upper-triangle storage with (anti)symmetric diagonal blocks (dbcsr_make_random_matrix,
    src/ops/dbcsr_test_methods.F:318-465)."""
    blocks = {}
    for r in range(1, len(row_sizes) + 1):
        for c in range(1, len(col_sizes) + 1):
            if symmetry != "N" and r > c:
                continue
            if rng.random() < sparsity:
                continue
            b = rng.uniform(-1.0, 1.0, (row_sizes[r - 1], col_sizes[c - 1]))
            if symmetry != "N" and r == c:
                b = 0.5 * (b + b.T) if symmetry == "S" else 0.5 * (b - b.T)
            blocks[(r, c)] = b
    return D.DbcsrMatrix.from_blocks(name, row_sizes, col_sizes, blocks, symmetry)


def check_multiply(matrix_c, dense_a, dense_b, dense_c_in, transa, transb, alpha, beta, limits, retain_sparsity):
    """dbcsr_check_multiply: returns eps_norm (<= 10 passes)."""
    dense_c_dbcsr = matrix_c.to_dense()
    m, n, k = limits[1] - limits[0] + 1, limits[3] - limits[2] + 1, limits[5] - limits[4] + 1
    r0, c0, k0 = limits[0] - 1, limits[2] - 1, limits[4] - 1
    a_sub = dense_a[r0:r0 + m, k0:k0 + k] if transa == "N" else dense_a[k0:k0 + k, r0:r0 + m].T
    b_sub = dense_b[k0:k0 + k, c0:c0 + n] if transb == "N" else dense_b[c0:c0 + n, k0:k0 + k].T
    dense_c = dense_c_in.copy()
    dense_c[r0:r0 + m, c0:c0 + n] = alpha * (a_sub @ b_sub) + beta * dense_c[r0:r0 + m, c0:c0 + n]
    if retain_sparsity:  # dbcsr_impose_sparsity: only the elements of C's (unchanged) block pattern
        mask = np.zeros_like(dense_c, dtype=bool)
        ro, co = matrix_c.row_blk_offset - 1, matrix_c.col_blk_offset - 1
        for (r, c) in matrix_c.blocks():
            mask[ro[r - 1]:ro[r], co[c - 1]:co[c]] = True
            if matrix_c.has_symmetry():
                mask[ro[c - 1]:ro[c], co[r - 1]:co[r]] = True
        dense_c[~mask] = 0.0
    inf_norm = lambda x: float(np.abs(x).sum(axis=1).max()) if x.size else 0.0  # noqa: E731  (dlange 'I')
    residual = inf_norm(dense_c - dense_c_dbcsr)
    eps = np.finfo(np.float64).eps / 2  # dlamch('eps') = relative machine epsilon 2^-53
    denom = (inf_norm(a_sub if transa == "N" else a_sub.T) + inf_norm(b_sub if transb == "N" else b_sub.T) + inf_norm(dense_c_in)) * n * eps
    return residual / denom if denom > 0 else (0.0 if residual == 0 else np.inf)


def run_case(case, backend, rng, symmetries=SYMMETRIES, transposes=("N", "T"), filter_eps=None):
    """dbcsr_test_multiplies for one parameter set: yields (description, eps_norm, flop) per multiply performed."""
    name, sizes, sparsities, retain, alpha, beta, bs_m, bs_n, bs_k, limits = case
    for a_symm, b_symm, c_symm in symmetries:
        if (a_symm != "N" or b_symm != "N") and sizes[0] != sizes[1]:
            continue
        if (a_symm != "N" or b_symm != "N") and sizes[0] != sizes[2]:
            continue
        if c_symm != "N" and sizes[0] != sizes[1]:
            continue
        for transa in transposes:
            for transb in transposes:
                if c_symm != "N":
                    if not ((transa == "N" and transb != "N") or (transa != "N" and transb == "N")):
                        continue
                    if limits[0] != 1 or limits[1] != sizes[0] or limits[2] != 1 or limits[3] != sizes[1]:
                        continue
                sizes_m = orc.random_block_sizes(sizes[0], bs_m)
                sizes_n = orc.random_block_sizes(sizes[1], bs_n)
                sizes_k = orc.random_block_sizes(sizes[2], bs_k)
                a_s, b_s, c_s = a_symm != "N", b_symm != "N", c_symm != "N"
                if (c_s and a_s and b_s) or (not c_s and a_s and b_s) or (c_s and not a_s and b_s) or (c_s and a_s and not b_s):
                    my_m, my_n, my_k = sizes_m, sizes_m, sizes_m
                elif not c_s and not a_s and b_s:
                    my_m, my_n, my_k = sizes_m, sizes_n, sizes_n
                elif not c_s and a_s and not b_s:
                    my_m, my_n, my_k = sizes_m, sizes_n, sizes_m
                elif c_s and not a_s and not b_s:
                    my_m, my_n, my_k = sizes_m, sizes_m, sizes_k
                else:
                    my_m, my_n, my_k = sizes_m, sizes_n, sizes_k
                matrix_c = random_matrix("Matrix C", my_m, my_n, sparsities[2], c_symm, rng)
                matrix_a = random_matrix("Matrix A", my_k if transa != "N" else my_m, my_m if transa != "N" else my_k, sparsities[0], a_symm, rng)
                matrix_b = random_matrix("Matrix B", my_n if transb != "N" else my_k, my_k if transb != "N" else my_n, sparsities[1], b_symm, rng)
                if c_symm != "N":
                    matrix_b = matrix_a.copy("Matrix B")
                    if c_symm == "A":
                        matrix_b.data *= -1.0
                dense_a, dense_b, dense_c = matrix_a.to_dense(), matrix_b.to_dense(), matrix_c.to_dense()
                # limits are given for the (m, n, k) of the case; block-size vectors were swapped for some symmetry mixes
                lim = list(limits)
                nfr, nfc, nfk = int(np.sum(my_m)), int(np.sum(my_n)), int(np.sum(my_k))
                lim[1], lim[3], lim[5] = min(lim[1], nfr), min(lim[3], nfc), min(lim[5], nfk)
                if lim[0] > lim[1] or lim[2] > lim[3] or lim[4] > lim[5]:
                    continue
                flop = D.dbcsr_multiply(transa, transb, alpha, matrix_a, matrix_b, beta, matrix_c, first_row=lim[0], last_row=lim[1],
                                        first_column=lim[2], last_column=lim[3], first_k=lim[4], last_k=lim[5], retain_sparsity=retain,
                                        filter_eps=filter_eps, backend=backend)
                eps_norm = check_multiply(matrix_c, dense_a, dense_b, dense_c, transa, transb, alpha, beta, lim, retain)
                yield "%s (%s,%s | %s,%s,%s)" % (name, transa, transb, a_symm, b_symm, c_symm), eps_norm, flop


# ---------------------------------------------------------------------------------------------------- golden .perf cases
def golden_cases():
    import json
    import os

    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "perf_golden.json")))["cases"]


def to_dbcsr(bm, name):
    """oracle BlockMatrix (BCSR-ordered block list, contiguous data) -> DbcsrMatrix sharing the data."""
    m = D.DbcsrMatrix(name, bm.row_blk_size, bm.col_blk_size)
    m.row_p = np.concatenate([[0], np.cumsum(np.bincount(bm.rows - 1, minlength=len(bm.row_blk_size)))])
    m.col_i = bm.cols.astype(np.int32)
    m.blk_p = (bm.offsets + 1).astype(np.int32)
    m.data = np.array(bm.data, dtype=np.float64, copy=True)
    return m


def run_golden_case(case, backend):
    """The reference's perf driver for one tests/inputs/*.perf file (tests/dbcsr_performance_multiply.F:271,373-411,623-677):
    matrices C, A, B from the dlarnv generator with counters 12341314/15/16, C := alpha op(A) op(B) + beta C through dbcsr_multiply,
    then dbcsr_checksum (plain and position-weighted).  Returns (checksum, checksum_pos)."""
    sizes_m = orc.random_block_sizes(case["M"], case["bs_m"])
    sizes_n = orc.random_block_sizes(case["N"], case["bs_n"])
    sizes_k = orc.random_block_sizes(case["K"], case["bs_k"])
    ta = case["transa"] == "T"
    C = orc.random_matrix(sizes_m, sizes_n, case["sparsity"][2], 12341314)
    A = orc.random_matrix(sizes_k, sizes_m, case["sparsity"][0], 12341315) if ta else orc.random_matrix(sizes_m, sizes_k, case["sparsity"][0], 12341315)
    B = orc.random_matrix(sizes_k, sizes_n, case["sparsity"][1], 12341316)
    mc = to_dbcsr(C, "C")
    D.dbcsr_multiply(case["transa"], case["transb"], case["alpha"][0], to_dbcsr(A, "A"), to_dbcsr(B, "B"), case["beta"][0], mc, backend=backend)
    out = orc.BlockMatrix(mc.row_blk_size, mc.col_blk_size, mc.block_rows(), mc.col_i, data=mc.data)
    assert np.array_equal(out.offsets + 1, mc.blk_p)  # finalized: compact data area in index order
    cs, cs_pos = out.checksum(), out.checksum(pos=True)
    # the product-side dbcsr_checksum agrees with the oracle's restatement
    assert abs(D.dbcsr_checksum(mc) / cs - 1.0) <= 1e-13 and abs(D.dbcsr_checksum(mc, pos=True) / cs_pos - 1.0) <= 1e-12
    return cs, cs_pos
