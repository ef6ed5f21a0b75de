"""Device-side stack builder (dbcsr_b200/csrc/host/device_builder.cu) against the host builder.

The device builder restates csr_multiply_low / flush_stacks / stack_sort as data-parallel passes (rank sorts, scans, atomic-min
first-touch table, stable radix sorts).  Here, on a machine without a GPU, the SAME pass functors run in host loops
(mode DEVICE_BUILD without LAUNCH); tests/test_gpu_device_builder.py runs them as CUDA kernels.  The bar is identity with the host
builder: the same stacks (7-wide entries in traversal order, 3-wide entries in device order), dispatched in the same order, the
same C index in first-touch order with the same offsets, the same flop count."""
import numpy as np
import pytest

from dbcsr_b200 import host
from test_host_builder import CASES, random_lists


def engines(m_sizes, n_sizes, k_sizes, nthreads, cfg_kw):
    ref = host.Engine(m_sizes, n_sizes, k_sizes, nthreads=nthreads, mode=host.RECORD, cfg=host.default_cfg(**cfg_kw))
    dev = host.Engine(m_sizes, n_sizes, k_sizes, nthreads=nthreads, mode=host.RECORD | host.DEVICE_BUILD, cfg=host.default_cfg(**cfg_kw))
    return ref, dev


def assert_same(ref, dev, nthreads):
    a, b = ref.stacks(), dev.stacks()
    assert len(a) == len(b)
    for x, y in zip(a, b):
        for key in ("m", "n", "k", "max_m", "max_n", "max_k", "defined_mnk", "stack_id", "thread"):
            assert x[key] == y[key], key
        assert np.array_equal(x["host"], y["host"])
        assert np.array_equal(x["dev"], y["dev"])
    for t in range(nthreads):
        for u, v in zip(ref.c_index(t), dev.c_index(t)):
            assert np.array_equal(u, v)
    assert ref.flop() == dev.flop()


@pytest.mark.parametrize("case", CASES, ids=[str(i) for i in range(len(CASES))])
def test_emulated_device_builder_matches_host_builder(case):
    nrow, ncol, nk, oa, ob, sizes, ssz, nst, lim = case
    m_sizes, n_sizes, k_sizes, A, B = random_lists(nrow, ncol, nk, oa, ob, sizes, seed=sum(case[:3]))
    a_l, b_l = np.array(A.index_list(), dtype=np.int32), np.array(B.index_list(), dtype=np.int32)
    ref, dev = engines(m_sizes, n_sizes, k_sizes, 1, dict(mm_stack_size=ssz, n_stacks=nst, multrec_limit=lim))
    ref.multiply(a_l, None, b_l, None)
    dev.multiply(a_l, None, b_l, None)
    assert len(ref.stacks()) > 0
    assert dev.device_built_ticks == 1 and ref.device_built_ticks == 0
    assert_same(ref, dev, 1)
    ref.close()
    dev.close()


@pytest.mark.parametrize("case", [CASES[1], CASES[4]], ids=["mixed", "23"])
def test_open_addressing_c_table(case, monkeypatch):
    """Block grids above the direct-table limit use an open-addressing (row, col) table that grows between ticks."""
    monkeypatch.setenv("DBCSR_B200_DEVBUILD_DENSE_LIMIT", "0")
    nrow, ncol, nk, oa, ob, sizes, ssz, nst, lim = case
    m_sizes, n_sizes, k_sizes, A, B = random_lists(nrow, ncol, nk, oa, ob, sizes, seed=77)
    a_l, b_l = np.array(A.index_list(), dtype=np.int32), np.array(B.index_list(), dtype=np.int32)
    ref, dev = engines(m_sizes, n_sizes, k_sizes, 1, dict(mm_stack_size=ssz, n_stacks=nst, multrec_limit=lim))
    for e in (ref, dev):
        e.multiply(a_l[: len(a_l) // 8], None, b_l, None)  # a small first tick, then ticks that force the table to grow
        e.multiply(a_l, None, b_l, None)
        e.multiply(a_l, None, b_l, None)
    assert dev.device_built_ticks == 3
    assert_same(ref, dev, 1)
    ref.close()
    dev.close()


@pytest.mark.parametrize("nthreads,row_chunks", [(1, 4), (3, 1), (2, 3)])
def test_threads_row_chunks_ticks_and_reset(nthreads, row_chunks):
    """Several host threads / row slices per thread (every slice ends with a purge), a second Cannon tick accumulating onto the
    index of the first (existing blocks keep their ids, new ones are appended), then a reset and a different product."""
    m_sizes, n_sizes, k_sizes, A, B = random_lists(90, 70, 80, 0.25, 0.2, [5, 13, 23], seed=11)
    _, _, _, A2, B2 = random_lists(90, 70, 80, 0.15, 0.3, [5, 13, 23], seed=12)
    A2 = type(A)(m_sizes, k_sizes, A2.rows, A2.cols)
    B2 = type(B)(k_sizes, n_sizes, B2.rows, B2.cols)
    kw = dict(mm_stack_size=400, n_stacks=3, multrec_limit=48, row_chunks=row_chunks)
    ref, dev = engines(m_sizes, n_sizes, k_sizes, nthreads, kw)
    for e in (ref, dev):
        e.multiply(np.array(A.index_list(), dtype=np.int32), None, np.array(B.index_list(), dtype=np.int32), None)
        e.multiply(np.array(A2.index_list(), dtype=np.int32), None, np.array(B2.index_list(), dtype=np.int32), None)
    assert_same(ref, dev, nthreads)
    for e in (ref, dev):
        e.reset()
        e.multiply(np.array(A2.index_list(), dtype=np.int32), None, np.array(B.index_list(), dtype=np.int32), None)
    assert_same(ref, dev, nthreads)
    ref.close()
    dev.close()


def test_no_sort_and_empty_panels():
    m_sizes, n_sizes, k_sizes, A, B = random_lists(40, 30, 50, 0.3, 0.3, [13, 23], seed=3)
    a_l, b_l = np.array(A.index_list(), dtype=np.int32), np.array(B.index_list(), dtype=np.int32)
    ref, dev = engines(m_sizes, n_sizes, k_sizes, 1, dict(mm_stack_size=250, stack_sort=0))
    for e in (ref, dev):
        e.multiply(a_l, None, b_l, None)
        e.multiply(a_l[:0], None, b_l, None)
        e.multiply(a_l, None, b_l[:0], None)
    assert_same(ref, dev, 1)
    ref.close()
    dev.close()


@pytest.mark.parametrize("keep", [False, True])
def test_preset_c_and_retain_sparsity(keep):
    """Existing C blocks (beta != 0) and retain_sparsity on the device passes: the work index starts from the listed blocks, with
    retain_sparsity products into other blocks are dropped and no block is created (src/mm/dbcsr_mm_csr.F:300-323, 540-576)."""
    m_sizes, n_sizes, k_sizes, A, B = random_lists(40, 36, 44, 0.25, 0.25, [5, 13, 23], seed=31)
    a_l, b_l = np.array(A.index_list(), dtype=np.int32), np.array(B.index_list(), dtype=np.int32)
    rng = np.random.default_rng(3)
    pres = np.nonzero(rng.random((40, 36)) < 0.3)
    rows, cols = (pres[0] + 1).astype(np.int32), (pres[1] + 1).astype(np.int32)
    for nthreads in (1, 3):
        ref, dev = engines(m_sizes, n_sizes, k_sizes, nthreads, dict(mm_stack_size=300, multrec_limit=64))
        for e in (ref, dev):
            e.preset_c(rows, cols, None, keep_sparsity=keep)
            e.multiply(a_l, None, b_l, None)
            e.multiply(a_l[::3], None, b_l, None)
        assert dev.device_built_ticks == 2 * nthreads
        assert_same(ref, dev, nthreads)
        for e in (ref, dev):  # both settings end with the multiply
            e.reset()
            e.multiply(a_l, None, b_l, None)
        assert_same(ref, dev, nthreads)
        ref.close()
        dev.close()


@pytest.mark.parametrize("eps", [1e-2, 1e-1, 1.0])
def test_on_the_fly_filter(eps):
    """filter_eps: a product is skipped when a_norm * b_norm < row_max_epss(row) in single precision (src/mm/dbcsr_mm_csr.F:270-278)."""
    from test_host_builder import _filter_case

    m_sizes, n_sizes, k_sizes, A, B, a_l, b_l, a_n, b_n, counts = _filter_case(seed=11)
    row_eps = host.row_max_epss(eps, counts)
    for nthreads, chunks in ((1, 1), (3, 2)):
        ref, dev = engines(m_sizes, n_sizes, k_sizes, nthreads, dict(mm_stack_size=400, multrec_limit=64, row_chunks=chunks))
        plain = host.Engine(m_sizes, n_sizes, k_sizes, nthreads=nthreads, mode=host.RECORD, cfg=host.default_cfg(mm_stack_size=400, multrec_limit=64, row_chunks=chunks))
        plain.multiply(a_l, None, b_l, None)
        for e in (ref, dev):
            e.set_filter(row_eps)
            e.multiply(a_l, None, b_l, None, a_norms=a_n, b_norms=b_n)
        assert dev.device_built_ticks == nthreads
        assert_same(ref, dev, nthreads)
        assert dev.flop() < plain.flop()
        for e in (ref, dev):  # filter off again
            e.reset()
            e.set_filter(None)
            e.multiply(a_l, None, b_l, None, a_norms=a_n, b_norms=b_n)
        assert_same(ref, dev, nthreads)
        assert dev.flop() == plain.flop()
        for e in (ref, dev, plain):
            e.close()


@pytest.mark.parametrize("use_maps", [False, True])
def test_symmetric_product_skipping(use_maps):
    """Product with symmetry: the half of the off-diagonal blocks the checkerboard rule stores transposed is skipped
    (src/mm/dbcsr_mm_csr.F:280-292), with and without local -> global index maps."""
    n, nk = 36, 40
    m_sizes, n_sizes, k_sizes, A, B = random_lists(n, n, nk, 0.3, 0.3, [5, 13, 23], seed=77)
    a_l, b_l = np.array(A.index_list(), dtype=np.int32), np.array(B.index_list(), dtype=np.int32)
    grows = gcols = None
    if use_maps:
        grows = 2 * np.arange(n, dtype=np.int32) + 1
        gcols = 2 * np.arange(n, dtype=np.int32) + 2
    for nthreads in (1, 4):
        ref, dev = engines(m_sizes, n_sizes, k_sizes, nthreads, dict(mm_stack_size=500, multrec_limit=64))
        for e in (ref, dev):
            e.set_c_symmetry(True, grows, gcols)
            e.multiply(a_l, None, b_l, None)
        assert dev.device_built_ticks == nthreads
        assert_same(ref, dev, nthreads)
        ref.close()
        dev.close()


def test_fuzz_device_builder_equals_host_builder():
    """Random shapes (incl. 1-sized dims and blocks), occupations, stack sizes, multrec limits, thread counts, row chunks, with
    existing C blocks, retain_sparsity, symmetric products and a second tick: identical stacks, device orders, C indices, flop."""
    rng = np.random.default_rng(4711)
    for it in range(30):
        nrow, ncol, nk = (int(x) for x in rng.integers(1, 60, 3))
        sizes = [int(x) for x in rng.choice([1, 2, 4, 5, 7, 13, 23, 26, 32, 40], size=int(rng.integers(1, 5)), replace=False)]
        m_sizes, n_sizes, k_sizes = rng.choice(sizes, nrow), rng.choice(sizes, ncol), rng.choice(sizes, nk)
        oa, ob = rng.uniform(0.05, 0.9, 2)
        ar, ac = np.nonzero(rng.random((nrow, nk)) < oa)
        br, bc = np.nonzero(rng.random((nk, ncol)) < ob)
        from oracle import oracle as orc

        A = orc.BlockMatrix(m_sizes, k_sizes, ar + 1, ac + 1)
        B = orc.BlockMatrix(k_sizes, n_sizes, br + 1, bc + 1)
        a_l = np.array(A.index_list(), dtype=np.int32).reshape(-1, 3)
        b_l = np.array(B.index_list(), dtype=np.int32).reshape(-1, 3)
        kw = dict(mm_stack_size=int(rng.choice([16, 50, 300, 30000])), n_stacks=int(rng.choice([3, 5])), multrec_limit=int(rng.choice([4, 32, 512])),
                  row_chunks=int(rng.choice([1, 2, 3])), stack_sort=int(rng.choice([1, 1, 0])))
        nthreads = int(rng.choice([1, 2, 3]))
        sym = bool(rng.random() < 0.3) and nrow == ncol
        keep = bool(rng.random() < 0.3)
        preset = bool(rng.random() < 0.5) or keep
        ref, dev = engines(m_sizes, n_sizes, k_sizes, nthreads, kw)
        if preset:
            pr, pc = (x + 1 for x in np.nonzero(rng.random((nrow, ncol)) < 0.3))
            for e in (ref, dev):
                e.preset_c(pr, pc, None, keep_sparsity=keep)
        if sym:
            for e in (ref, dev):
                e.set_c_symmetry(True)
        for tick in range(2):
            for e in (ref, dev):
                e.multiply(a_l, None, b_l, None)
        assert dev.device_built_ticks == 2 * nthreads, it
        assert_same(ref, dev, nthreads)
        ref.close()
        dev.close()


@pytest.mark.parametrize("nthreads,row_chunks,tile", [(1, 1, 8), (2, 3, 16), (1, 2, 1000)])
def test_tile_order_same_products_same_index(nthreads, row_chunks, tile):
    """dev_tile > 0: NOT the reference's stacks, but the same C index (first-touch order, offsets), the same multiset of
    (a, b, c) entries per thread and stack shape, stacks of at most mm_stack_size entries, and inside a stack every C block in
    one run of consecutive entries."""
    m_sizes, n_sizes, k_sizes, A, B = random_lists(70, 60, 80, 0.3, 0.3, [23], seed=31)
    a_l, b_l = np.array(A.index_list(), dtype=np.int32), np.array(B.index_list(), dtype=np.int32)
    kw = dict(mm_stack_size=500, multrec_limit=64, row_chunks=row_chunks)
    ref = host.Engine(m_sizes, n_sizes, k_sizes, nthreads=nthreads, mode=host.RECORD, cfg=host.default_cfg(**kw))
    dev = host.Engine(m_sizes, n_sizes, k_sizes, nthreads=nthreads, mode=host.RECORD | host.DEVICE_BUILD, cfg=host.default_cfg(dev_tile=tile, **kw))
    for e in (ref, dev):
        e.multiply(a_l, None, b_l, None)
        e.multiply(a_l[::2], None, b_l, None)  # second tick onto the same index
    assert dev.device_built_ticks == 2 * nthreads
    for t in range(nthreads):
        for u, v in zip(ref.c_index(t), dev.c_index(t)):
            assert np.array_equal(u, v)
    assert ref.flop() == dev.flop()
    for t in range(nthreads):
        ea = np.concatenate([s["host"] for s in ref.stacks() if s["thread"] == t])
        eb = np.concatenate([s["host"] for s in dev.stacks() if s["thread"] == t])
        assert np.array_equal(ea[np.lexsort(ea.T[::-1])], eb[np.lexsort(eb.T[::-1])])
    nruns_ref = nruns_dev = 0
    for s in dev.stacks():
        assert 0 < s["dev"].shape[0] <= 500 and s["defined_mnk"]
        assert np.array_equal(s["dev"], s["host"][:, 3:6])  # a tile-ordered stack is handed over in its device order
        c = s["dev"][:, 2]
        starts = np.flatnonzero(np.r_[True, c[1:] != c[:-1]])
        assert len(set(c[starts].tolist())) == starts.size  # every C block in ONE run
        nruns_dev += starts.size
    for s in ref.stacks():
        c = s["dev"][:, 2]
        nruns_ref += 1 + int(np.count_nonzero(c[1:] != c[:-1]))
    assert nruns_dev <= nruns_ref
    ref.close()
    dev.close()


@pytest.mark.parametrize("cfg_name,nblk,n_stacks", [("cfg2", 300, 3), ("cfg3", 240, 5), ("cfg2", None, 3), ("cfg3", None, 5)],
                         ids=["cfg2_300", "cfg3_240", "cfg2_full_size", "cfg3_full_size"])
def test_baseline_like_workloads_default_stack_size(cfg_name, nblk, n_stacks):
    """BASELINE workloads (23x23 at 10 %, mixed {5,13,23,26,32} at 5 %) on reduced grids AND at their full 1000 x 1000 size (1e7 / 2.5e6
    products, 337 / ~130 stacks) with the DEFAULT 30000-entry stacks, two threads x four row slices: fill events inside slices, purges
    at slice ends, binned (< 4000 flop) and sorted stacks side by side -- identical to the host builder."""
    from dbcsr_b200 import workload

    w = workload.make_config(cfg_name, nblk=nblk)
    A, B, bs = w["A"], w["B"], w["m_sizes"]
    kw = dict(n_stacks=n_stacks, row_chunks=4)
    ref, dev = engines(bs, bs, bs, 2, kw)
    for e in (ref, dev):
        e.multiply(A.list3(), None, B.list3(), None)
    assert dev.device_built_ticks == 2 and len(ref.stacks()) > 8
    assert_same(ref, dev, 2)
    ref.close()
    dev.close()
