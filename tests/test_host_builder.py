"""CPU tests of the host-side stack builder (dbcsr_b200/csrc/host, C++) against the pure-Python index oracle
(oracle/index_oracle.py): identical rec_sort order, identical stacks (contents + dispatch order), identical C block index
(order of first touch = "bit-identical block index structure"), identical flop count."""
import numpy as np
import pytest

from dbcsr_b200 import host
from oracle import index_oracle as io
from oracle import oracle as orc


def random_lists(nrow, ncol, nk, occ_a, occ_b, sizes, seed):
    rng = np.random.default_rng(seed)
    m_sizes = rng.choice(sizes, nrow)
    n_sizes = rng.choice(sizes, ncol)
    k_sizes = rng.choice(sizes, nk)
    A = orc.BlockMatrix(m_sizes, k_sizes, *np.nonzero(rng.random((nrow, nk)) < occ_a))
    A.rows += 1
    A.cols += 1
    B = orc.BlockMatrix(k_sizes, n_sizes, *np.nonzero(rng.random((nk, ncol)) < occ_b))
    B.rows += 1
    B.cols += 1
    A = orc.BlockMatrix(m_sizes, k_sizes, A.rows, A.cols)
    B = orc.BlockMatrix(k_sizes, n_sizes, B.rows, B.cols)
    return m_sizes, n_sizes, k_sizes, A, B


def test_rec_sort_index_matches_oracle():
    rng = np.random.default_rng(0)
    for (nr, nc, nb) in [(1, 1, 1), (7, 3, 12), (50, 60, 400), (64, 64, 1000), (100, 17, 900), (3, 200, 300)]:
        cells = rng.choice(nr * nc, size=min(nb, nr * nc), replace=False)
        cells.sort()
        lst = [(int(c // nc) + 1, int(c % nc) + 1, i + 1) for i, c in enumerate(cells)]
        exp = io.rec_sort_index(1, nr, 1, nc, list(lst)) if len(lst) > 1 else lst
        got = host.rec_sort_index(nr, nc, np.array(lst, dtype=np.int32))
        assert [tuple(int(v) for v in r) for r in got] == exp


def test_stack_sort_and_binning_match_oracle():
    rng = np.random.default_rng(1)
    S = 5000
    p = np.zeros((S, 7), dtype=np.int32)
    p[:, 3:6] = rng.integers(1, 10000, (S, 3))
    p[:, 5] = rng.integers(0, 300, S) * 25 + 1
    o = io.LocalMultiplyOracle([5], [5], [5])
    assert np.array_equal(host.stack_sort(p), np.array(o._stack_sort([tuple(r) for r in p.tolist()]), dtype=np.int32))
    assert np.array_equal(host.stack_binning(p), np.array(o._stack_binning([tuple(r) for r in p.tolist()]), dtype=np.int32))
    # offsets beyond 46339: the reference's 32-bit product val(3)*(val(3)+3) wraps before it is widened (accdrv.F:405-406)
    p[:, 5] = rng.integers(0, 400000, S) * 25 + 1
    assert np.array_equal(host.stack_binning(p), np.array(o._stack_binning([tuple(r) for r in p.tolist()]), dtype=np.int32))
    c = 50000
    assert ((c * (c + 3)) + 2 ** 31) % 2 ** 32 - 2 ** 31 < 0  # the wrapped product is negative here; MODULO keeps the bin id >= 0


CASES = [
    # nrow, ncol, nk, occA, occB, sizes, stack_size, n_stacks, multrec_limit
    (40, 40, 40, 0.3, 0.3, [23], 1000, 3, 512),
    (60, 50, 70, 0.2, 0.25, [5, 13, 23, 26, 32], 300, 3, 64),
    (60, 50, 70, 0.2, 0.25, [5, 13, 23, 26, 32], 300, 5, 64),
    (30, 100, 20, 0.5, 0.1, [4, 5, 7], 200, 3, 32),
    (128, 128, 128, 0.1, 0.1, [23], 4000, 3, 512),
    (10, 10, 10, 1.0, 1.0, [1, 3, 4], 50, 3, 8),
    (33, 1, 17, 0.6, 0.9, [5, 8, 9], 64, 3, 16),
    (30, 5000, 40, 0.3, 0.004, [5, 13], 300, 3, 64),  # more than 4096 block columns: per-row hash tables instead of direct tables
]


@pytest.mark.parametrize("case", CASES, ids=[str(i) for i in range(len(CASES))])
def test_engine_matches_index_oracle(case):
    nrow, ncol, nk, oa, ob, sizes, ssz, nst, lim = case
    m_sizes, n_sizes, k_sizes, A, B = random_lists(nrow, ncol, nk, oa, ob, sizes, seed=sum(case[:3]))
    ora = io.LocalMultiplyOracle(m_sizes, n_sizes, k_sizes, mm_stack_size=ssz, n_stacks=nst, multrec_limit=lim)
    exp = ora.multiply(A.index_list(), B.index_list())
    eng = host.Engine(m_sizes, n_sizes, k_sizes, nthreads=1, mode=host.RECORD,
                      cfg=host.default_cfg(mm_stack_size=ssz, n_stacks=nst, multrec_limit=lim))
    eng.multiply(np.array(A.index_list(), dtype=np.int32), None, np.array(B.index_list(), dtype=np.int32), None)
    got = eng.stacks()
    assert len(got) == len(exp)
    for g, x in zip(got, exp):
        for key in ("m", "n", "k", "max_m", "max_n", "max_k", "defined_mnk", "stack_id"):
            assert g[key] == x[key], key
        assert np.array_equal(g["host"], x["host"])
        assert np.array_equal(g["dev"], x["dev"])
    rows, cols, blk_p, datasize = eng.c_index(0)
    assert list(rows) == ora.c_row_i and list(cols) == ora.c_col_i and list(blk_p) == ora.c_blk_p
    assert datasize == ora.datasize and eng.flop() == ora.flop
    # every (A blk, B blk) pair with matching k appears exactly once over all stacks
    n_products = sum(g["host"].shape[0] for g in got)
    kb = {}
    for r in B.rows:
        kb[int(r)] = kb.get(int(r), 0) + 1
    assert n_products == sum(kb.get(int(c), 0) for c in A.cols)
    eng.close()


def test_multithreaded_engine_same_products_disjoint_rows():
    """T threads: same multiset of (a,b) products, C rows disjoint between threads, per-thread C index self-consistent."""
    m_sizes, n_sizes, k_sizes, A, B = random_lists(96, 80, 64, 0.2, 0.2, [5, 13, 23], seed=5)
    a_l, b_l = np.array(A.index_list(), dtype=np.int32), np.array(B.index_list(), dtype=np.int32)
    e1 = host.Engine(m_sizes, n_sizes, k_sizes, nthreads=1, cfg=host.default_cfg(mm_stack_size=500))
    e1.multiply(a_l, None, b_l, None)
    e4 = host.Engine(m_sizes, n_sizes, k_sizes, nthreads=4, cfg=host.default_cfg(mm_stack_size=500))
    e4.multiply(a_l, None, b_l, None)
    prod = lambda eng: sorted((int(r[3]), int(r[4])) for s in eng.stacks() for r in s["host"])
    assert prod(e1) == prod(e4) and e1.flop() == e4.flop()
    seen_rows = set()
    for t in range(4):
        rows, cols, blk_p, ds = e4.c_index(t)
        assert not (set(rows.tolist()) & seen_rows)
        seen_rows |= set(rows.tolist())
        sizes = m_sizes[rows - 1] * n_sizes[cols - 1]
        assert np.array_equal(blk_p, 1 + np.concatenate([[0], np.cumsum(sizes)[:-1]])) and ds == int(sizes.sum())
    e1.close()
    e4.close()


def test_edge_cases_empty_and_ragged():
    """Empty panels, a single block, more threads than block rows, no matching k, repeated multiplies with reset-free engines."""
    bs = np.array([23, 5, 13, 23], dtype=np.int32)
    empty = np.zeros((0, 3), dtype=np.int32)
    one_a = np.array([[2, 3, 1]], dtype=np.int32)
    one_b = np.array([[3, 1, 1]], dtype=np.int32)
    for nthreads in (1, 3, 9):
        e = host.Engine(bs, bs, bs, nthreads=nthreads, mode=host.RECORD)
        e.multiply(empty, None, empty, None)
        assert e.stacks() == [] and e.flop() == 0
        e.multiply(one_a, None, empty, None)
        e.multiply(empty, None, one_b, None)
        assert e.stacks() == []
        e.multiply(one_a, None, one_b, None)  # A(2,3) * B(3,1) -> C(2,1): m=5, k=13, n=23
        st = e.stacks()
        assert len(st) == 1 and st[0]["host"].tolist() == [[5, 23, 13, 1, 1, 1, 1]]
        assert e.flop() == 2 * 5 * 23 * 13
        # no common k: A(1,2) with B(3,1)
        e.multiply(np.array([[1, 2, 1]], dtype=np.int32), None, one_b, None)
        assert len(e.stacks()) == 1
        e.close()


def test_second_tick_accumulates_into_existing_c_blocks():
    """Two ticks on one engine (Cannon): C blocks touched again keep their offset, new ones are appended (first touch order)."""
    bs = np.full(6, 23, dtype=np.int32)
    a1 = np.array([[1, 1, 1], [2, 2, 530]], dtype=np.int32)
    b1 = np.array([[1, 1, 1], [2, 1, 530]], dtype=np.int32)
    e = host.Engine(bs, bs, bs, nthreads=1, mode=host.RECORD)
    e.multiply(a1, None, b1, None)   # C(1,1), C(2,1)
    a2 = np.array([[1, 3, 1], [3, 3, 530]], dtype=np.int32)
    b2 = np.array([[3, 1, 1], [3, 2, 530]], dtype=np.int32)
    e.multiply(a2, None, b2, None)   # C(1,1) again, C(1,2), C(3,1), C(3,2) new
    rows, cols, blk_p, ds = e.c_index(0)
    got = list(zip(rows.tolist(), cols.tolist(), blk_p.tolist()))
    assert got[:2] == [(1, 1, 1), (2, 1, 530)]
    assert sorted(got[2:]) == [(1, 2, 1059), (3, 1, 1588), (3, 2, 2117)] or len(got) == 5
    assert ds == 5 * 529 and len(set((r, c) for r, c, _ in got)) == 5
    # the repeated C(1,1) product points at offset 1
    second = e.stacks()[-1]["host"]
    assert any(row[5] == 1 for row in second.tolist())
    e.close()


def test_row_chunks_same_products_and_row_ordered_c():
    """row_chunks > 1: every thread walks several block-row chunks in row order; products unchanged, C rows disjoint between
    threads, and inside a thread the C blocks of a chunk are contiguous (what the early per-chunk D2H relies on)."""
    m_sizes, n_sizes, k_sizes, A, B = random_lists(90, 70, 60, 0.2, 0.2, [5, 13, 23], seed=9)
    a_l, b_l = np.array(A.index_list(), dtype=np.int32), np.array(B.index_list(), dtype=np.int32)
    e1 = host.Engine(m_sizes, n_sizes, k_sizes, nthreads=1, cfg=host.default_cfg(mm_stack_size=300))
    e1.multiply(a_l, None, b_l, None)
    ec = host.Engine(m_sizes, n_sizes, k_sizes, nthreads=2, cfg=host.default_cfg(mm_stack_size=300, row_chunks=3))
    ec.multiply(a_l, None, b_l, None)
    prod = lambda eng: sorted((int(r[3]), int(r[4])) for s in eng.stacks() for r in s["host"])
    assert prod(e1) == prod(ec) and e1.flop() == ec.flop()
    seen = set()
    for t in range(2):
        rows, cols, blk_p, ds = ec.c_index(t)
        assert not (set(rows.tolist()) & seen)
        seen |= set(rows.tolist())
        chunk_of = (rows - 1) * 6 // 90  # chunk id of every C block (6 chunks over 90 rows)
        assert np.all(np.diff(chunk_of) >= 0)          # chunk by chunk
        assert set(chunk_of.tolist()) <= {t, t + 2, t + 4}  # thread t owns chunks t, t+2, t+4
    e1.close()
    ec.close()


MIXES = [[23], [5, 13, 23, 26, 32], [1, 3, 4], [4, 5, 7], [5, 8, 9], [4, 13, 25], [14, 29, 32], [45, 67, 78]]


@pytest.mark.parametrize("sizes", MIXES, ids=[str(s) for s in MIXES])
def test_recorded_host_stacks_reproduce_the_block_product(sizes):
    """CPU end-to-end of the host side: stacks built by the C++ engine (block-size mixes of tests/dbcsr_unittest3.F:76-118, default
    N_STACKS=3 so that inhomogeneous default stacks occur) and drained with the oracle's blas_process_mm_stack restatement give the
    same C as the oracle's plain block product, and the same (row, col) structure."""
    rng = np.random.default_rng(len(sizes) * 11 + sizes[0])
    nr, nc, nk = 30, 26, 34
    ms, ns, ks = (rng.choice(sizes, n).astype(np.int32) for n in (nr, nc, nk))
    amask, bmask = rng.random((nr, nk)) < 0.3, rng.random((nk, nc)) < 0.3
    ar, ac = np.nonzero(amask)
    br, bc = np.nonzero(bmask)
    A = orc.BlockMatrix(ms, ks, ar + 1, ac + 1)
    B = orc.BlockMatrix(ks, ns, br + 1, bc + 1)
    A.data[:] = rng.random(A.nze)
    B.data[:] = rng.random(B.nze)
    eng = host.Engine(ms, ns, ks, nthreads=2, mode=host.RECORD, cfg=host.default_cfg(mm_stack_size=150))
    eng.multiply(np.array(A.index_list(), dtype=np.int32), None, np.array(B.index_list(), dtype=np.int32), None)
    got = {}
    for t in range(2):
        rows, cols, blk_p, ds = eng.c_index(t)
        c = np.zeros(max(ds, 1))
        for st in eng.stacks():
            if st["thread"] == t:
                orc.host_stack(st["host"], A.data, B.data, c)
        for r, cc, o in zip(rows, cols, blk_p):
            got[(int(r), int(cc))] = c[o - 1:o - 1 + int(ms[r - 1]) * int(ns[cc - 1])]
    Cref = orc.multiply_blocks(A, B)
    assert set(got) == set(zip(Cref.rows.tolist(), Cref.cols.tolist()))
    num = den = 0.0
    for r, cc, o in zip(Cref.rows, Cref.cols, Cref.offsets):
        ref = Cref.data[o:o + int(ms[r - 1]) * int(ns[cc - 1])]
        num += float(((got[(int(r), int(cc))] - ref) ** 2).sum())
        den += float((ref ** 2).sum())
    assert (num / den) ** 0.5 <= 1e-13
    eng.close()


# ---------------------------------------------------------------------------------------------- on-the-fly norm filter
def _filter_case(seed, nrow=48, ncol=40, nk=56, sizes=(5, 13, 23)):
    m_sizes, n_sizes, k_sizes, A, B = random_lists(nrow, ncol, nk, 0.3, 0.3, list(sizes), seed=seed)
    rng = np.random.default_rng(seed + 100)
    a_l, b_l = np.array(A.index_list(), dtype=np.int32), np.array(B.index_list(), dtype=np.int32)
    # squared block norms spread over 8 decades so that a mid-range eps removes a good part of the products
    a_n = (10.0 ** rng.uniform(-8, 0, a_l.shape[0])).astype(np.float32)
    b_n = (10.0 ** rng.uniform(-8, 0, b_l.shape[0])).astype(np.float32)
    counts = np.bincount(a_l[:, 0] - 1, minlength=nrow)
    return m_sizes, n_sizes, k_sizes, A, B, a_l, b_l, a_n, b_n, counts


def test_row_max_epss_matches_oracle_single_precision():
    counts = np.array([0, 1, 2, 3, 7, 100, 12345], dtype=np.int32)
    for eps in (1e-5, 1e-7, 3.3e-10, 0.0):
        got = host.row_max_epss(eps, counts)
        assert got.dtype == np.float32 and np.array_equal(got, io.row_max_epss(eps, counts))


@pytest.mark.parametrize("eps", [1e-2, 1e-1, 1.0])
def test_filtered_engine_matches_index_oracle(eps):
    """filter_eps: product skipped when a_norm*b_norm < row_max_epss(row) (src/mm/dbcsr_mm_csr.F:270-278); same stacks and
    the same C index as the restated reference traversal, and strictly fewer products than without the filter."""
    m_sizes, n_sizes, k_sizes, A, B, a_l, b_l, a_n, b_n, counts = _filter_case(seed=11)
    row_eps = io.row_max_epss(eps, counts)
    ora = io.LocalMultiplyOracle(m_sizes, n_sizes, k_sizes, mm_stack_size=400, multrec_limit=64)
    exp = ora.multiply(A.index_list(), B.index_list(), a_norms=a_n, b_norms=b_n, row_eps=row_eps)
    assert ora.skipped > 0
    eng = host.Engine(m_sizes, n_sizes, k_sizes, nthreads=1, mode=host.RECORD, cfg=host.default_cfg(mm_stack_size=400, multrec_limit=64))
    eng.set_filter(host.row_max_epss(eps, counts))
    eng.multiply(a_l, None, b_l, None, a_norms=a_n, b_norms=b_n)
    got = eng.stacks()
    assert len(got) == len(exp)
    for g, x in zip(got, exp):
        assert g["stack_id"] == x["stack_id"] and np.array_equal(g["host"], x["host"]) and np.array_equal(g["dev"], x["dev"])
    rows, cols, blk_p, datasize = eng.c_index(0)
    assert list(rows) == ora.c_row_i and list(cols) == ora.c_col_i and list(blk_p) == ora.c_blk_p and datasize == ora.datasize
    assert eng.flop() == ora.flop
    # unfiltered run on the same engine type has more products
    e0 = host.Engine(m_sizes, n_sizes, k_sizes, nthreads=1, mode=host.RECORD, cfg=host.default_cfg(mm_stack_size=400, multrec_limit=64))
    e0.multiply(a_l, None, b_l, None)
    assert e0.flop() > eng.flop()
    n_kept = sum(g["host"].shape[0] for g in got)
    assert n_kept + ora.skipped == sum(s["host"].shape[0] for s in e0.stacks())
    e0.close()
    eng.close()


def test_filter_keeps_exactly_the_products_above_threshold_multithreaded():
    """Independent of the traversal: the set of surviving (a_blk, b_blk) pairs is {a_norm*b_norm >= row_eps(row)} in float32,
    for 1 and 4 threads and with row chunks; filter off (set_filter(None) or eps = 0) gives the unfiltered product set."""
    m_sizes, n_sizes, k_sizes, A, B, a_l, b_l, a_n, b_n, counts = _filter_case(seed=23, nrow=64)
    row_eps = host.row_max_epss(0.2, counts)
    by_k = {}
    for j in range(b_l.shape[0]):
        by_k.setdefault(int(b_l[j, 0]), []).append(j)
    want, everything = set(), set()
    for i in range(a_l.shape[0]):
        for j in by_k.get(int(a_l[i, 1]), []):
            everything.add((int(a_l[i, 2]), int(b_l[j, 2])))
            if not (np.float32(a_n[i] * b_n[j]) < row_eps[a_l[i, 0] - 1]):
                want.add((int(a_l[i, 2]), int(b_l[j, 2])))
    assert 0 < len(want) < len(everything)
    prod = lambda eng: sorted((int(r[3]), int(r[4])) for s in eng.stacks() for r in s["host"])
    for nthreads, chunks in ((1, 1), (4, 1), (3, 2)):
        e = host.Engine(m_sizes, n_sizes, k_sizes, nthreads=nthreads, mode=host.RECORD,
                        cfg=host.default_cfg(mm_stack_size=300, row_chunks=chunks))
        e.set_filter(row_eps)
        e.multiply(a_l, None, b_l, None, a_norms=a_n, b_norms=b_n)
        assert prod(e) == sorted(want)
        e.reset()
        e.set_filter(None)
        e.multiply(a_l, None, b_l, None, a_norms=a_n, b_norms=b_n)
        assert prod(e) == sorted(everything)
        e.reset()
        e.set_filter(host.row_max_epss(0.0, counts))
        e.multiply(a_l, None, b_l, None, a_norms=a_n, b_norms=b_n)
        assert prod(e) == sorted(everything)
        e.close()


# ---------------------------------------------------------------------------------------------- beta != 0, retain_sparsity, final filter
@pytest.mark.parametrize("keep", [False, True])
def test_preset_c_and_retain_sparsity_match_index_oracle(keep):
    """Work matrix starting from existing C blocks (beta != 0) and retain_sparsity: same stacks and C index as the restatement
    of fill_hash_tables + dbcsr_mm_csr_multiply_low (src/mm/dbcsr_mm_csr.F:300-323,540-576)."""
    m_sizes, n_sizes, k_sizes, A, B = random_lists(40, 36, 44, 0.25, 0.25, [5, 13, 23], seed=31)
    rng = np.random.default_rng(3)
    pres = np.nonzero(rng.random((40, 36)) < 0.3)
    c_rows, c_cols = pres[0] + 1, pres[1] + 1
    ora = io.LocalMultiplyOracle(m_sizes, n_sizes, k_sizes, mm_stack_size=300, multrec_limit=64)
    ora.preset_c(c_rows, c_cols, keep_sparsity=keep)
    exp = ora.multiply(A.index_list(), B.index_list())
    eng = host.Engine(m_sizes, n_sizes, k_sizes, nthreads=1, mode=host.RECORD, cfg=host.default_cfg(mm_stack_size=300, multrec_limit=64))
    eng.preset_c(c_rows, c_cols, keep_sparsity=keep)
    eng.multiply(np.array(A.index_list(), dtype=np.int32), None, np.array(B.index_list(), dtype=np.int32), None)
    got = eng.stacks()
    assert len(got) == len(exp) and len(got) > 0
    for g, x in zip(got, exp):
        assert g["stack_id"] == x["stack_id"] and np.array_equal(g["host"], x["host"]) and np.array_equal(g["dev"], x["dev"])
    rows, cols, blk_p, datasize = eng.c_index(0)
    assert list(rows) == ora.c_row_i and list(cols) == ora.c_col_i and list(blk_p) == ora.c_blk_p and datasize == ora.datasize
    assert eng.flop() == ora.flop
    if keep:  # no new blocks, every product lands in a listed block
        assert rows.size == c_rows.size
        listed = set(zip(c_rows.tolist(), c_cols.tolist()))
        assert all((int(rows[r[6] - 1]), int(cols[r[6] - 1])) in listed for s in got for r in s["host"])
    else:
        assert rows.size > c_rows.size and list(rows[:c_rows.size]) == c_rows.tolist()
    with pytest.raises(Exception):  # only before the first tick
        eng.preset_c(c_rows, c_cols)
    eng.reset()  # both settings end with the multiply
    eng.multiply(np.array(A.index_list(), dtype=np.int32), None, np.array(B.index_list(), dtype=np.int32), None)
    ora2 = io.LocalMultiplyOracle(m_sizes, n_sizes, k_sizes, mm_stack_size=300, multrec_limit=64)
    ora2.multiply(A.index_list(), B.index_list())
    assert list(eng.c_index(0)[0]) == ora2.c_row_i and eng.flop() == ora2.flop
    eng.close()


def test_preset_c_multithreaded_blocks_go_to_row_owners():
    m_sizes, n_sizes, k_sizes, A, B = random_lists(50, 30, 40, 0.25, 0.25, [5, 13], seed=41)
    rng = np.random.default_rng(4)
    pres = np.nonzero(rng.random((50, 30)) < 0.4)
    perm = rng.permutation(pres[0].size)  # arbitrary list order
    c_rows, c_cols = pres[0][perm] + 1, pres[1][perm] + 1
    a_l, b_l = np.array(A.index_list(), dtype=np.int32), np.array(B.index_list(), dtype=np.int32)
    e1 = host.Engine(m_sizes, n_sizes, k_sizes, nthreads=1, cfg=host.default_cfg(mm_stack_size=300))
    e1.preset_c(c_rows, c_cols, keep_sparsity=True)
    e1.multiply(a_l, None, b_l, None)
    e3 = host.Engine(m_sizes, n_sizes, k_sizes, nthreads=3, cfg=host.default_cfg(mm_stack_size=300, row_chunks=2))
    e3.preset_c(c_rows, c_cols, keep_sparsity=True)
    e3.multiply(a_l, None, b_l, None)
    prod = lambda eng: sorted((int(r[3]), int(r[4])) for s in eng.stacks() for r in s["host"])
    assert prod(e1) == prod(e3) and e1.flop() == e3.flop()
    seen, total = set(), 0
    for t in range(3):
        rows, cols, blk_p, ds = e3.c_index(t)
        total += rows.size
        assert not (set(rows.tolist()) & seen)
        seen |= set(rows.tolist())
        sizes = m_sizes[rows - 1] * n_sizes[cols - 1]
        assert np.array_equal(blk_p, 1 + np.concatenate([[0], np.cumsum(sizes)[:-1]])) and ds == int(sizes.sum())
    assert total == c_rows.size
    e1.close()
    e3.close()


def test_filter_index_matches_multrec_filtering_restatement():
    rng = np.random.default_rng(8)
    rbs, cbs = rng.choice([5, 13, 23], 30), rng.choice([5, 13, 23], 25)
    rows, cols = (x + 1 for x in np.nonzero(rng.random((30, 25)) < 0.5))
    perm = rng.permutation(rows.size)
    rows, cols = rows[perm], cols[perm]
    nel = rbs[rows - 1] * cbs[cols - 1]
    blk_p = 1 + np.concatenate([[0], np.cumsum(nel)[:-1]])
    data = rng.standard_normal(int(nel.sum()))
    for i in range(rows.size):
        data[blk_p[i] - 1:blk_p[i] - 1 + nel[i]] *= 10.0 ** rng.uniform(-6, 0)
    blk_p[5] = 0  # deleted block: skipped
    for eps in (0.0, 1e-4, 1e-2, 1.0, 1e3):
        er, ec, ep, enze, norms = io.multrec_filtering(eps, rows, cols, blk_p, rbs, cbs, data)
        gr, gc, gp, gnze = host.filter_index(eps, norms, rows, cols, blk_p, nel)
        assert gr.tolist() == er and gc.tolist() == ec and gp.tolist() == ep and gnze == enze
    assert len(io.multrec_filtering(1e-2, rows, cols, blk_p, rbs, cbs, data)[0]) not in (0, rows.size - 1)


@pytest.mark.parametrize("use_maps", [False, True])
def test_symmetric_product_skipping_matches_index_oracle(use_maps):
    """Product matrix with symmetry (src/mm/dbcsr_mm_csr.F:280-292): only the blocks for which the checkerboard rule
    checker_tr(global row, global col) is false are computed (plus the diagonal); stacks, C index and flop equal the oracle's,
    for 1 and 4 threads the surviving product set is exactly {(i,j): i == j or not checker_tr(i,j)}."""
    n, nk = 36, 40
    m_sizes, n_sizes, k_sizes, A, B = random_lists(n, n, nk, 0.3, 0.3, [5, 13, 23], seed=77)
    rng = np.random.default_rng(3)
    grows = gcols = None
    if use_maps:  # local -> global block indices of a 2-column process grid (every second global row/col is local)
        grows = 2 * np.arange(n, dtype=np.int32) + 1
        gcols = 2 * np.arange(n, dtype=np.int32) + 2
    ora = io.LocalMultiplyOracle(m_sizes, n_sizes, k_sizes, mm_stack_size=500, n_stacks=3, multrec_limit=64)
    ora.set_c_symmetry(True, grows, gcols)
    exp = ora.multiply(A.index_list(), B.index_list())
    cfg = host.default_cfg(mm_stack_size=500, n_stacks=3, multrec_limit=64)
    eng = host.Engine(m_sizes, n_sizes, k_sizes, nthreads=1, mode=host.RECORD, cfg=cfg)
    eng.set_c_symmetry(True, grows, gcols)
    eng.multiply(np.array(A.index_list(), dtype=np.int32), None, np.array(B.index_list(), dtype=np.int32), None)
    got = eng.stacks()
    assert len(got) == len(exp) and len(exp) > 0
    for g, x in zip(got, exp):
        assert np.array_equal(g["host"], x["host"]) and np.array_equal(g["dev"], x["dev"])
    rows, cols, blk_p, datasize = eng.c_index(0)
    assert list(rows) == ora.c_row_i and list(cols) == ora.c_col_i and list(blk_p) == ora.c_blk_p
    assert eng.flop() == ora.flop
    gr = np.arange(1, n + 1) if grows is None else grows
    gc = np.arange(1, n + 1) if gcols is None else gcols
    for r, c in zip(rows, cols):
        assert gr[r - 1] == gc[c - 1] or not io.checker_tr(int(gr[r - 1]), int(gc[c - 1]))
    # the skipped half is really skipped, the kept half complete: compare with the unrestricted product pattern
    full = host.Engine(m_sizes, n_sizes, k_sizes, nthreads=1, mode=host.RECORD, cfg=cfg)
    full.multiply(np.array(A.index_list(), dtype=np.int32), None, np.array(B.index_list(), dtype=np.int32), None)
    fr, fc, _, _ = full.c_index(0)
    want = {(int(r), int(c)) for r, c in zip(fr, fc) if gr[r - 1] == gc[c - 1] or not io.checker_tr(int(gr[r - 1]), int(gc[c - 1]))}
    assert {(int(r), int(c)) for r, c in zip(rows, cols)} == want
    eng4 = host.Engine(m_sizes, n_sizes, k_sizes, nthreads=4, mode=host.RECORD, cfg=cfg)
    eng4.set_c_symmetry(True, grows, gcols)
    eng4.multiply(np.array(A.index_list(), dtype=np.int32), None, np.array(B.index_list(), dtype=np.int32), None)
    got4 = set()
    for t in range(4):
        r4, c4, _, _ = eng4.c_index(t)
        got4 |= {(int(r), int(c)) for r, c in zip(r4, c4)}
    assert got4 == want and eng4.flop() == eng.flop()
    # the setting ends with reset(): the next multiply computes everything again
    eng.reset()
    eng.multiply(np.array(A.index_list(), dtype=np.int32), None, np.array(B.index_list(), dtype=np.int32), None)
    assert eng.flop() == full.flop()
    for e in (eng, full, eng4):
        e.close()


def test_fuzz_engine_equals_index_oracle_incl_presets_symmetry_and_second_tick():
    """Random shapes (incl. 1-sized dims and blocks), occupations, stack sizes, multrec limits, with existing C blocks,
    retain_sparsity, symmetric products, and a second Cannon tick on the same C: the C++ builder (direct row tables, counting /
    radix ordering by block id) and the Python restatement produce identical stacks, device orders, C indices and flop counts."""
    rng = np.random.default_rng(2024)
    for it in range(25):
        nrow, ncol, nk = (int(x) for x in rng.integers(1, 60, 3))
        sizes = [int(x) for x in rng.choice([1, 2, 4, 5, 7, 13, 23, 26, 32, 40], size=int(rng.integers(1, 5)), replace=False)]
        m_sizes, n_sizes, k_sizes = rng.choice(sizes, nrow), rng.choice(sizes, ncol), rng.choice(sizes, nk)
        oa, ob = rng.uniform(0.05, 0.9, 2)
        A = orc.BlockMatrix(m_sizes, k_sizes, *(x + 1 for x in np.nonzero(rng.random((nrow, nk)) < oa)))
        B = orc.BlockMatrix(k_sizes, n_sizes, *(x + 1 for x in np.nonzero(rng.random((nk, ncol)) < ob)))
        ssz, nst, lim = int(rng.choice([16, 50, 300, 30000])), int(rng.choice([3, 5])), int(rng.choice([4, 32, 512]))
        sym = bool(rng.random() < 0.3) and nrow == ncol
        keep = bool(rng.random() < 0.3)
        preset = bool(rng.random() < 0.5) or keep
        ora = io.LocalMultiplyOracle(m_sizes, n_sizes, k_sizes, mm_stack_size=ssz, n_stacks=nst, multrec_limit=lim)
        eng = host.Engine(m_sizes, n_sizes, k_sizes, nthreads=1, mode=host.RECORD,
                          cfg=host.default_cfg(mm_stack_size=ssz, n_stacks=nst, multrec_limit=lim))
        if preset:
            pr, pc = (x + 1 for x in np.nonzero(rng.random((nrow, ncol)) < 0.3))
            ora.preset_c(pr, pc, keep_sparsity=keep)
            eng.preset_c(pr, pc, None, keep_sparsity=keep)
        if sym:
            ora.set_c_symmetry(True)
            eng.set_c_symmetry(True)
        a_l = np.array(A.index_list(), dtype=np.int32).reshape(-1, 3)
        b_l = np.array(B.index_list(), dtype=np.int32).reshape(-1, 3)
        for tick in range(2):
            exp = ora.multiply(A.index_list(), B.index_list())
            eng.multiply(a_l, None, b_l, None)
            got = eng.stacks()
            assert len(got) == len(exp), (it, tick)
            for g, x in zip(got, exp):
                assert np.array_equal(g["host"], x["host"]) and np.array_equal(g["dev"], x["dev"]), (it, tick)
        rows, cols, blk_p, ds = eng.c_index(0)
        assert list(rows) == ora.c_row_i and list(cols) == ora.c_col_i and list(blk_p) == ora.c_blk_p and ds == ora.datasize
        assert eng.flop() == ora.flop
        eng.close()


@pytest.mark.parametrize("nthreads", [1, 4])
def test_host_only_recorder_library_matches_the_engine(nthreads):
    """libdbcsr_b200_hostbuilder.so (no accelerator code; what bench.py's CPU reference arm loads) gives the engine's stacks."""
    from dbcsr_b200 import hostbuilder

    m_sizes, n_sizes, k_sizes, A, B = random_lists(70, 60, 50, 0.25, 0.25, [5, 13, 23, 26, 32], seed=3)
    a_l, b_l = np.array(A.index_list(), dtype=np.int32), np.array(B.index_list(), dtype=np.int32)
    eng = host.Engine(m_sizes, n_sizes, k_sizes, nthreads=nthreads, mode=host.RECORD, cfg=host.default_cfg(mm_stack_size=300, n_stacks=3))
    eng.multiply(a_l, None, b_l, None)
    ref = eng.stacks()
    got, datasizes, flop = hostbuilder.record_stacks(m_sizes, n_sizes, k_sizes, a_l, b_l, nthreads=nthreads, mm_stack_size=300, n_stacks=3)
    assert len(got) == len(ref) > 0 and flop == eng.flop()
    for g, x in zip(got, ref):
        for key in ("m", "n", "k", "defined_mnk", "thread", "stack_id"):
            assert g[key] == x[key], key
        assert np.array_equal(g["host"], x["host"])
    assert datasizes == [eng.c_index(t)[3] for t in range(nthreads)]
    eng.close()
    import subprocess

    lib = hostbuilder._lib()._name
    deps = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
    assert "cuda" not in deps.lower() and "dbcsr_acc" not in deps
