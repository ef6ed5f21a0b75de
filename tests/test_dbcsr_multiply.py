"""CPU tests of the dbcsr_multiply mirror (dbcsr_b200/dbcsr.py): the complete pre/post-processing of the operator (transposes,
symmetry expansion, limits / cropping, alpha / beta, retain_sparsity, symmetric products in checkerboard positions, finalize)
runs unchanged; only the local multiply is served by the ORACLE (index oracle builds the stacks in the reference's order, the C
restatement of blas_process_mm_stack_d drains them) because this container has no GPU.  The GPU twin of this file
(tests/test_gpu_dbcsr_multiply.py) runs the same cases through the device engine."""
import zlib

import numpy as np
import pytest

from dbcsr_b200 import dbcsr as D
from oracle import index_oracle as io
from oracle import oracle as orc

from dbcsr_multiply_cases import SYMMETRIES, UNITTEST1_CASES, check_multiply, golden_cases, random_matrix, run_case, run_golden_case


class OracleBackend:
    """Test-only stand-in for DeviceBackend.local_multiply (same contract)."""

    def local_multiply(self, m_sizes, n_sizes, k_sizes, left, right, c_preset, keep_sparsity, c_symmetry, filter_eps, final_filter):
        ora = io.LocalMultiplyOracle(m_sizes, n_sizes, k_sizes, mm_stack_size=1000, n_stacks=3, multrec_limit=512)
        c_data = np.zeros(0)
        if c_preset is not None:
            ora.preset_c(c_preset[0], c_preset[1], keep_sparsity=keep_sparsity)
            c_data = np.array(c_preset[2], dtype=np.float64)
        if c_symmetry:
            ora.set_c_symmetry(True)
        a_list = [tuple(int(v) for v in r) for r in left.list3()]
        b_list = [tuple(int(v) for v in r) for r in right.list3()]
        kw = {}
        if filter_eps is not None:
            kw["a_norms"] = orc.norms(left.data, left.blk_p - 1, m_sizes[left.block_rows() - 1] * k_sizes[left.col_i - 1])
            kw["b_norms"] = orc.norms(right.data, right.blk_p - 1, k_sizes[right.block_rows() - 1] * n_sizes[right.col_i - 1])
            counts = np.bincount(left.block_rows() - 1, minlength=len(m_sizes))
            kw["row_eps"] = (np.float32(filter_eps) / np.maximum(1, counts).astype(np.float32)) ** 2
        stacks = ora.multiply(a_list, b_list, **kw)
        c = np.zeros(ora.datasize)
        c[:c_data.size] = c_data
        for st in stacks:
            orc.host_stack(st["host"], left.data, right.data, c)
        rows, cols, blk_p = np.array(ora.c_row_i, dtype=np.int32), np.array(ora.c_col_i, dtype=np.int32), np.array(ora.c_blk_p, dtype=np.int32)
        if final_filter:  # multrec_filtering, src/mm/dbcsr_mm_multrec.F:700-758
            nze = np.asarray(m_sizes)[rows - 1].astype(np.int64) * np.asarray(n_sizes)[cols - 1]
            keep = np.array([nze[i] > 0 and float(np.dot(c[blk_p[i] - 1:blk_p[i] - 1 + nze[i]], c[blk_p[i] - 1:blk_p[i] - 1 + nze[i]])) >= filter_eps ** 2
                             for i in range(rows.size)], dtype=bool)
            rows, cols, blk_p = rows[keep], cols[keep], blk_p[keep]
        return [(rows, cols, blk_p, c)], ora.flop


@pytest.mark.parametrize("case", UNITTEST1_CASES, ids=[c[0] for c in UNITTEST1_CASES])
def test_dbcsr_multiply_unittest_cases(case):
    rng = np.random.default_rng(zlib.crc32(case[0].encode()))  # deterministic per case
    n = 0
    for desc, eps_norm, flop in run_case(case, OracleBackend(), rng):
        assert eps_norm <= 10.0, (desc, eps_norm)
        n += 1
    assert n >= 1, n


def test_abort_messages_match_reference():
    rng = np.random.default_rng(1)
    a = random_matrix("A", [2, 3], [4, 1], 0.0, "N", rng)
    b = random_matrix("B", [4, 1], [3, 3], 0.0, "N", rng)
    c = random_matrix("C", [2, 3], [3, 3], 0.0, "N", rng)
    be = OracleBackend()
    for kw, msg in [(dict(first_row=0), "Invalid first row specified"), (dict(last_row=6), "Invalid last row specified"),
                    (dict(first_column=7), "Invalid first col specified"), (dict(last_column=7), "Invalid last column specified (C)"),
                    (dict(first_k=6), "Invalid first k specified (A)"), (dict(last_k=6), "Invalid last k specified (A)"),
                    (dict(first_row=4, last_row=2), "Last row smaller than first row"),
                    (dict(first_column=5, last_column=2), "Last col smaller than first col")]:
        with pytest.raises(D.DbcsrAbort, match=msg.replace("(", r"\(").replace(")", r"\)")):
            D.dbcsr_multiply("N", "N", 1.0, a, b, 0.0, c.copy(), backend=be, **kw)
    with pytest.raises(D.DbcsrAbort, match="wrong transa_l = X"):
        D.dbcsr_multiply("x", "N", 1.0, a, b, 0.0, c.copy(), backend=be)
    with pytest.raises(D.DbcsrAbort, match="wrong transb_l = Q"):
        D.dbcsr_multiply("N", "q", 1.0, a, b, 0.0, c.copy(), backend=be)
    with pytest.raises(D.DbcsrAbort, match="C/A rows not equal"):
        D.dbcsr_multiply("T", "N", 1.0, a, b, 0.0, c.copy(), backend=be)
    with pytest.raises(D.DbcsrAbort, match="A cols/B rows not equal"):
        D.dbcsr_multiply("N", "N", 1.0, a, random_matrix("B", [3, 2], [3, 3], 0.0, "N", rng), 0.0, c.copy(), backend=be)
    with pytest.raises(D.DbcsrAbort, match="C/B columns not equal"):
        D.dbcsr_multiply("N", "N", 1.0, a, random_matrix("B", [4, 1], [2, 4], 0.0, "N", rng), 0.0, c.copy(), backend=be)


def test_filter_eps_drops_small_blocks_and_keeps_the_rest_accurate():
    """filter_eps: on-the-fly filter + final filter (src/mm/dbcsr_mm.F docs :363-374): every surviving block has norm >= eps, every
    dropped block of the exact product has norm < eps * (a slack for the products skipped on the fly)."""
    rng = np.random.default_rng(5)
    sizes = orc.random_block_sizes(92, [1, 5, 1, 13, 1, 23])
    a = random_matrix("A", sizes, sizes, 0.6, "N", rng)
    b = random_matrix("B", sizes, sizes, 0.6, "N", rng)
    # scale some blocks down so that the filter has something to do
    for i in range(0, a.nblks, 3):
        a.block(i)[...] *= 1e-7
    exact = a.to_dense() @ b.to_dense()
    c = D.DbcsrMatrix("C", sizes, sizes)
    eps = 1e-4
    D.dbcsr_multiply("N", "N", 1.0, a, b, 0.0, c, filter_eps=eps, backend=OracleBackend())
    ro, co = c.row_blk_offset - 1, c.col_blk_offset - 1
    got = c.blocks()
    assert 0 < len(got) < len(sizes) ** 2
    for (r, cc), blk in got.items():
        assert np.linalg.norm(blk) >= eps
        assert np.abs(blk - exact[ro[r - 1]:ro[r], co[cc - 1]:co[cc]]).max() <= 1e-3 * eps * len(sizes) + 1e-12
    for r in range(1, len(sizes) + 1):
        for cc in range(1, len(sizes) + 1):
            if (r, cc) not in got:
                assert np.linalg.norm(exact[ro[r - 1]:ro[r], co[cc - 1]:co[cc]]) < 2 * eps


def test_finalize_merges_work_matrices_into_bcsr():
    """dbcsr_finalize: per-thread work matrices in first-touch order -> one BCSR index (rows ascending, cols sorted), data compact."""
    rs, cs = [2, 3, 1], [1, 4]
    p0 = (np.array([3, 1, 1]), np.array([2, 2, 1]), np.array([1, 5, 13]), np.arange(1.0, 15.0))      # rows 3,1,1
    p1 = (np.array([2]), np.array([1]), np.array([1]), np.array([7.0, 8.0, 9.0]))
    m = D.dbcsr_finalize(rs, cs, [p0, p1])
    assert m.row_p.tolist() == [0, 2, 3, 4] and m.col_i.tolist() == [1, 2, 1, 2]
    assert m.blk_p.tolist() == [1, 3, 11, 14]
    b = m.blocks()
    assert b[(1, 1)].T.reshape(-1).tolist() == [13.0, 14.0] and b[(2, 1)].T.reshape(-1).tolist() == [7.0, 8.0, 9.0]
    assert b[(3, 2)].T.reshape(-1).tolist() == [1.0, 2.0, 3.0, 4.0] and b[(1, 2)].shape == (2, 4)
    with pytest.raises(D.DbcsrAbort):
        D.dbcsr_finalize(rs, cs, [p0, (np.array([3]), np.array([2]), np.array([1]), np.zeros(4))])


def test_finalize_index_cpp_matches_python_finalize():
    """dbcsr_b200_finalize_index (C++, used by the device-side finalize) against dbcsr_finalize (Python) on random work indices."""
    from dbcsr_b200 import host

    rng = np.random.default_rng(0)
    for nr, nc, nb in [(20, 15, 120), (1, 1, 1), (7, 9, 0), (50, 3, 150)]:
        rs, cs = rng.integers(0, 6, nr), rng.integers(1, 6, nc)
        cells = rng.choice(nr * nc, nb, replace=False)
        rows, cols = (cells // nc + 1).astype(np.int32), (cells % nc + 1).astype(np.int32)
        ne = (rs[rows - 1] * cs[cols - 1]).astype(np.int32)
        blk_p = (1 + np.concatenate([[0], np.cumsum(ne)[:-1]])).astype(np.int32) if nb else np.zeros(0, dtype=np.int32)
        data = rng.random(int(ne.sum()))
        r, c, bp, perm, nze = host.finalize_index(rows, cols, ne)
        m = D.dbcsr_finalize(rs, cs, [(rows, cols, blk_p, data)])
        assert np.array_equal(c, m.col_i) and np.array_equal(bp, m.blk_p) and nze == data.size
        assert np.array_equal(r, m.block_rows())
        gathered = np.concatenate([data[blk_p[o] - 1:blk_p[o] - 1 + ne[o]] for o in perm]) if nb else np.zeros(0)
        assert np.array_equal(gathered, m.data)
    with pytest.raises(Exception):
        host.finalize_index([1, 1], [2, 2], [4, 4])


@pytest.mark.parametrize("case", golden_cases(), ids=[c["name"] for c in golden_cases()])
def test_perf_golden_checksums_through_dbcsr_multiply(case):
    """The operator mirror against the reference's STORED results: the nine golden checksum pairs of tests/inputs/*.perf
    (tests/golden/perf_golden.json) are reproduced when the multiply goes through dbcsr_multiply (transposes, beta = 1 keeps the
    existing C blocks, finalize) with the index oracle building the stacks -- i.e. operator logic + traversal + stack semantics
    are pinned end to end to numbers the reference itself produced (rel. 1e-11, tests/dbcsr_performance_multiply.F:656-677)."""
    cs, cs_pos = run_golden_case(case, OracleBackend())
    thr = max(case["threshold"], 1e-11)
    assert abs(cs / case["checksum"] - 1.0) <= thr, (cs, case["checksum"])
    assert abs(cs_pos / case["checksum_pos"] - 1.0) <= thr, (cs_pos, case["checksum_pos"])
