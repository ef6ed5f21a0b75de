"""Device-side stack builder on the GPU (dbcsr_b200/csrc/host/device_builder.cu).

(1) identity with the host builder: stacks (7-wide in traversal order, 3-wide in device order), dispatch order, C index, flop --
    the same cases as tests/test_device_builder_cpu.py, now with the passes running as CUDA kernels (cub scans / radix sorts);
(2) the products computed from device-built stacks against the oracle (block structure + values, north_star tolerance 1e-10)."""
import numpy as np
import pytest

from dbcsr_b200 import host, workload
from dbcsr_b200.multiply import DeviceMultiply
from test_gpu_multiply import check_against_oracle
from test_host_builder import CASES, random_lists

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def acc():
    from dbcsr_b200 import lib as acclib

    a = acclib.Acc(0)
    yield a
    a.finalize()


def same_stacks(ref, dev, nthreads):
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


def panels_from_lists(A, B, rng):
    """workload panels with random data for the block lists of the index cases"""
    def mk(M):
        order = np.lexsort((M.cols, M.rows))  # BCSR order
        return workload.Panel(M.row_blk_size, M.col_blk_size, np.asarray(M.rows)[order], np.asarray(M.cols)[order], rng=rng)

    return mk(A), mk(B)


@pytest.mark.parametrize("dense_table", [True, False], ids=["direct_table", "open_addressing"])
@pytest.mark.parametrize("case", CASES, ids=[str(i) for i in range(len(CASES))])
def test_device_builder_identical_to_host_builder(acc, case, dense_table, monkeypatch):
    if not dense_table:
        monkeypatch.setenv("DBCSR_B200_DEVBUILD_DENSE_LIMIT", "0")
    nrow, ncol, nk, oa, ob, sizes, ssz, nst, lim = case
    m_sizes, n_sizes, k_sizes, A, B = random_lists(nrow, ncol, nk, oa, ob, sizes, seed=sum(case[:3]))
    rng = np.random.default_rng(5)
    PA, PB = panels_from_lists(A, B, rng)
    cfg = dict(mm_stack_size=ssz, n_stacks=nst, multrec_limit=lim)
    ref = host.Engine(m_sizes, n_sizes, k_sizes, nthreads=1, mode=host.RECORD, cfg=host.default_cfg(**cfg))
    ref.multiply(PA.list3(), None, PB.list3(), None)
    dm = DeviceMultiply(acc, m_sizes, n_sizes, k_sizes, PA.data.size, PB.data.size, PB.nblks, nthreads=1, cfg=host.default_cfg(**cfg),
                        mode=host.LAUNCH | host.RECORD | host.DEVICE_BUILD)
    try:
        dm.upload_panels(PA.data, PB.data, PB.list3())
        dm.multiply(PA.list3(), PB.list3())
        prod = dm.download_c()
        assert dm.engine.device_built_ticks == 1
        same_stacks(ref, dm.engine, 1)
        check_against_oracle(PA, PB, prod, m_sizes, n_sizes)
    finally:
        dm.close()
        ref.close()


@pytest.mark.parametrize("nthreads,row_chunks", [(1, 4), (3, 2)])
def test_device_builder_threads_chunks_reset_and_early_download(acc, nthreads, row_chunks):
    """Row slices with their own purge and early D2H on the copy stream, pooled engine reused for a second multiply."""
    rng = np.random.default_rng(21)
    ms, ns, ks = (workload.block_sizes(n, [5, 13, 23], rng) for n in (96, 80, 72))
    cfg = dict(mm_stack_size=700, n_stacks=3, multrec_limit=64, row_chunks=row_chunks)
    dm = DeviceMultiply(acc, ms, ns, ks, 96 * 72 * 23 * 23, 72 * 80 * 23 * 23, 72 * 80, nthreads=nthreads, cfg=host.default_cfg(**cfg),
                        mode=host.LAUNCH | host.RECORD | host.DEVICE_BUILD)
    ref = host.Engine(ms, ns, ks, nthreads=nthreads, mode=host.RECORD, cfg=host.default_cfg(**cfg))
    try:
        bufs = None
        for rep in range(2):
            A = workload.random_panel(ms, ks, 0.25 + 0.1 * rep, rng)
            B = workload.random_panel(ks, ns, 0.3, rng)
            if rep:
                ref.reset()
            ref.multiply(A.list3(), None, B.list3(), None)
            dm.upload_panels(A.data, B.data, B.list3())
            dm.multiply(A.list3(), B.list3())
            if bufs is None:
                dm.engine.sync()
                bufs = [np.zeros(max(dm.engine.c_capacity(t), 1)) for t in range(nthreads)]
                prod = dm.download_c(bufs)
                dm.set_result_buffers(bufs)
            else:
                prod = dm.download_c()  # copies enqueued by the engine behind every slice
            assert dm.engine.device_built_ticks == nthreads * (rep + 1)
            same_stacks(ref, dm.engine, nthreads)
            check_against_oracle(A, B, prod, ms, ns)
    finally:
        dm.close()
        ref.close()


def test_device_builder_full_size_properties(acc):
    """BASELINE config 2 at full size through the device builder: flop, the C index of both host threads identical to the host
    builder's, and the sum property sum(C) == sum over products of colsum(A) . rowsum(B)."""
    w = workload.make_config("cfg2")
    A, B, bs = w["A"], w["B"], w["m_sizes"]
    cfg = dict(row_chunks=4)
    dm = DeviceMultiply(acc, bs, bs, bs, A.data.size, B.data.size, B.nblks, nthreads=2, cfg=host.default_cfg(**cfg),
                        mode=host.LAUNCH | host.DEVICE_BUILD)
    ref = host.Engine(bs, bs, bs, nthreads=2, mode=0, cfg=host.default_cfg(**cfg))
    try:
        ref.multiply(A.list3(), None, B.list3(), None)
        dm.upload_panels(A.data, B.data, B.list3())
        dm.multiply(A.list3(), B.list3())
        prod = dm.download_c()
        assert dm.engine.device_built_ticks == 2
        assert dm.engine.flop() == ref.flop()
        for t in range(2):
            for u, v in zip(ref.c_index(t), dm.engine.c_index(t)):
                assert np.array_equal(u, v)
        colsum_a = [A.block(i).sum(axis=0) for i in range(A.nblks)]
        rowsum_b = {}
        for i in range(B.nblks):
            rowsum_b.setdefault(int(B.rows[i]), []).append(B.block(i).sum(axis=1))
        rb = {k: np.sum(v, axis=0) for k, v in rowsum_b.items()}
        expected = sum(float(colsum_a[i] @ rb[int(A.cols[i])]) for i in range(A.nblks) if int(A.cols[i]) in rb)
        got = sum(float(p[3].sum()) for p in prod.parts)
        assert abs(got / expected - 1.0) <= 1e-10, (got, expected)
    finally:
        dm.close()
        ref.close()


@pytest.mark.parametrize("tile", [4, 64])
def test_tile_order_product_matches_oracle(acc, tile):
    """cfg.dev_tile: the device builder's tile order (same products and C index, other stack contents) gives the oracle's product."""
    rng = np.random.default_rng(8)
    ms, ns, ks = (workload.block_sizes(n, [23], rng) for n in (60, 70, 50))
    A = workload.random_panel(ms, ks, 0.3, rng)
    B = workload.random_panel(ks, ns, 0.3, rng)
    cfg = dict(mm_stack_size=900, multrec_limit=64, row_chunks=2)
    dm = DeviceMultiply(acc, ms, ns, ks, A.data.size, B.data.size, B.nblks, nthreads=2, cfg=host.default_cfg(dev_tile=tile, **cfg),
                        mode=host.LAUNCH | host.RECORD | host.DEVICE_BUILD)
    ref = host.Engine(ms, ns, ks, nthreads=2, mode=host.RECORD, cfg=host.default_cfg(**cfg))
    try:
        ref.multiply(A.list3(), None, B.list3(), None)
        dm.upload_panels(A.data, B.data, B.list3())
        dm.multiply(A.list3(), B.list3())
        prod = dm.download_c()
        assert dm.engine.device_built_ticks == 2
        for t in range(2):
            for u, v in zip(ref.c_index(t), dm.engine.c_index(t)):
                assert np.array_equal(u, v)
            ea = np.concatenate([s["host"] for s in ref.stacks() if s["thread"] == t])
            eb = np.concatenate([s["host"] for s in dm.engine.stacks() if s["thread"] == t])
            assert np.array_equal(ea[np.lexsort(ea.T[::-1])], eb[np.lexsort(eb.T[::-1])])
        for s in dm.engine.stacks():
            c = s["dev"][:, 2]
            starts = np.flatnonzero(np.r_[True, c[1:] != c[:-1]])
            assert len(set(c[starts].tolist())) == starts.size  # every C block in one run
        check_against_oracle(A, B, prod, ms, ns)
    finally:
        dm.close()
        ref.close()


FEATURES = [dict(preset=True), dict(preset=True, keep=True), dict(filter=0.05), dict(sym=True), dict(preset=True, filter=0.05, sym=True)]


@pytest.mark.parametrize("feat", FEATURES, ids=["beta", "retain_sparsity", "filter", "symmetry", "beta+filter+symmetry"])
@pytest.mark.parametrize("nthreads", [1, 3])
def test_device_builder_with_presets_filter_symmetry(acc, feat, nthreads):
    """beta != 0 / retain_sparsity / on-the-fly filter / symmetric product through the device passes: the same multiply with the host
    builder and with the device builder on the GPU gives identical stacks, dispatch order and C index, and the same C data."""
    rng = np.random.default_rng(13)
    n = 40
    ms = workload.block_sizes(n, [5, 13, 23], rng)
    ns, ks = ms, workload.block_sizes(44, [5, 13, 23], rng)
    A = workload.random_panel(ms, ks, 0.25, rng)
    B = workload.random_panel(ks, ns, 0.25, rng)
    for P in (A, B):  # block magnitudes over four decades so that the filter cuts a good part of the products
        for i in range(P.nblks):
            P.block(i)[...] *= 10.0 ** rng.uniform(-4, 0)
    C0 = workload.random_panel(ms, ns, 0.3, rng)
    results = []
    for mode in (host.LAUNCH | host.RECORD, host.LAUNCH | host.RECORD | host.DEVICE_BUILD):
        dm = DeviceMultiply(acc, ms, ns, ks, A.data.size, B.data.size, B.nblks, nthreads=nthreads,
                            cfg=host.default_cfg(mm_stack_size=400, multrec_limit=64, row_chunks=2), mode=mode)
        try:
            dm.upload_panels(A.data, B.data, B.list3())
            dm.multiply(A.list3(), B.list3(), filter_eps=feat.get("filter"),
                        c_preset=(C0.rows, C0.cols, 0.5 * C0.data) if feat.get("preset") else None,
                        retain_sparsity=bool(feat.get("keep")), c_symmetry=bool(feat.get("sym")))
            prod = dm.download_c()
            results.append((dm.engine.stacks(), [dm.engine.c_index(t) for t in range(nthreads)], dm.engine.flop(), prod.blocks(),
                            dm.engine.device_built_ticks))
        finally:
            dm.close()
    (st_h, idx_h, flop_h, blk_h, ticks_h), (st_d, idx_d, flop_d, blk_d, ticks_d) = results
    assert ticks_h == 0 and ticks_d == nthreads
    assert flop_h == flop_d and len(st_h) == len(st_d) and len(st_h) > 0
    for x, y in zip(st_h, st_d):
        assert x["stack_id"] == y["stack_id"] and x["thread"] == y["thread"]
        assert np.array_equal(x["host"], y["host"]) and np.array_equal(x["dev"], y["dev"])
    for a, b in zip(idx_h, idx_d):
        for u, v in zip(a, b):
            assert np.array_equal(u, v)
    assert set(blk_h) == set(blk_d)
    num = sum(float(((blk_h[k] - blk_d[k]) ** 2).sum()) for k in blk_h)
    den = sum(float((blk_h[k] ** 2).sum()) for k in blk_h)
    assert np.sqrt(num / max(den, 1e-300)) <= 1e-12
