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


def test_other_multiplies_use_the_host_builder():
    """Existing C blocks / retain_sparsity are outside the device passes: such a multiply must still give the host builder's result."""
    m_sizes, n_sizes, k_sizes, A, B = random_lists(30, 30, 30, 0.3, 0.3, [5, 13], seed=4)
    a_l, b_l = np.array(A.index_list(), dtype=np.int32), np.array(B.index_list(), dtype=np.int32)
    ref, dev = engines(m_sizes, n_sizes, k_sizes, 1, dict(mm_stack_size=300))
    rows, cols = np.array([1, 2, 3], dtype=np.int32), np.array([3, 2, 1], dtype=np.int32)
    for e in (ref, dev):
        e.preset_c(rows, cols, None, keep_sparsity=True)
        e.multiply(a_l, None, b_l, None)
    assert dev.device_built_ticks == 0
    assert_same(ref, dev, 1)
    for e in (ref, dev):  # and the next plain multiply is on the device passes again
        e.reset()
        e.multiply(a_l, None, b_l, None)
    assert_same(ref, dev, 1)
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
