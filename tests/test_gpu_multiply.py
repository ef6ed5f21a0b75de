"""End-to-end GPU parity of the local multiply (host panels -> C ABI -> C blocks on the host) against the oracle, modelled on
the reference's tests/dbcsr_test_multiply.F:523-759 (random sparse A, B -> reference product -> normwise criterion :753-759)
and tests/dbcsr_unittest3.F:76-118 (GPU-targeted block-size mixes).  Full BASELINE size is checked through size-independent
properties (sum of all C elements = sum over products of colsum(A).rowsum(B); C index = boolean product of the patterns)."""
import numpy as np
import pytest

from dbcsr_b200 import host, workload
from dbcsr_b200.multiply import DeviceMultiply, multiply
from oracle import index_oracle as io
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def acc():
    from dbcsr_b200 import lib as acclib

    a = acclib.Acc(0)
    yield a
    a.finalize()


def to_oracle(P):
    return orc.BlockMatrix(P.row_sizes, P.col_sizes, P.rows, P.cols, data=P.data)


def check_against_oracle(A, B, prod, m_sizes, n_sizes):
    Cref = orc.multiply_blocks(to_oracle(A), to_oracle(B))
    got = prod.blocks()
    ref_keys = set(zip(Cref.rows.tolist(), Cref.cols.tolist()))
    assert set(got.keys()) == ref_keys  # identical block structure
    num = den = 0.0
    amax = np.abs(A.data).max() if A.data.size else 0.0
    bmax = np.abs(B.data).max() if B.data.size else 0.0
    cmax = 0.0
    worst = 0.0
    for (r, c, o) in zip(Cref.rows, Cref.cols, Cref.offsets):
        m, n = int(m_sizes[r - 1]), int(n_sizes[c - 1])
        ref = Cref.data[o:o + m * n].reshape(n, m).T
        d = got[(int(r), int(c))] - ref
        num += float((d * d).sum())
        den += float((ref * ref).sum())
        cmax = max(cmax, float(np.abs(ref).max()))
        worst = max(worst, float(np.abs(d).max()))
    rel = np.sqrt(num / den) if den > 0 else 0.0
    assert rel <= 1e-10, rel  # north_star FP64 tolerance
    # tests/dbcsr_test_multiply.F:753-759: |C - C_ref|_inf / ((|A|+|B|+|C|) * N * eps) <= 10  (element-wise max norms here)
    N = int(np.sum(A.col_sizes))
    assert worst / ((amax + bmax + cmax) * N * np.finfo(np.float64).eps) <= 10.0
    # canonical BCSR index identical to the reference structure
    row_p, col_i = prod.bcsr_index()
    keys = sorted(ref_keys)
    assert list(col_i) == [c for _, c in keys]
    return rel


# block-size mixes of tests/dbcsr_unittest3.F:76-118 plus the BASELINE ones
MIXES = [[23], [5, 13, 23, 26, 32], [1, 3, 4], [4, 5, 7], [5, 8, 9], [4, 13, 25], [14, 29, 32], [45, 67, 78]]


@pytest.mark.parametrize("sizes", MIXES, ids=[str(s) for s in MIXES])
@pytest.mark.parametrize("nthreads", [1, 3])
def test_multiply_block_mixes(acc, sizes, nthreads):
    rng = np.random.default_rng(len(sizes) * 7 + sizes[0])
    nr, nc, nk = 48, 40, 56
    ms, ns, ks = (workload.block_sizes(n, sizes, rng) for n in (nr, nc, nk))
    A = workload.random_panel(ms, ks, 0.3, rng)
    B = workload.random_panel(ks, ns, 0.3, rng)
    # n_stacks=3 => at most 3 sizes per dimension are homogeneous; the rest lands in the inhomogeneous default stack,
    # which this library (like the reference, libsmm_acc.cpp:327) rejects with -1 => use enough stacks here
    cfg = host.default_cfg(mm_stack_size=500, n_stacks=max(3, len(sizes)))
    prod, flop = multiply(acc, A, B, ms, ns, ks, nthreads=nthreads, cfg=cfg)
    assert flop == int(sum(2 * int(ms[r - 1]) * int(ks[c - 1]) * int((ns[B.cols[B.rows == c] - 1]).sum()) for r, c in zip(A.rows, A.cols)))
    check_against_oracle(A, B, prod, ms, ns)


@pytest.mark.parametrize("sizes,n_stacks", [([5, 13, 23, 26, 32], 3), ([4, 7, 9, 45, 67], 2), ([3, 100, 23], 1)])
def test_inhomogeneous_stacks_run_on_the_gpu(acc, sizes, n_stacks):
    """More block sizes than DBCSR_N_STACKS: the rest lands in the inhomogeneous default stack (def_mnk = 0).  The reference
    returns -1 for it (CPU fall-back); this library bins it by shape and drains it on the GPU (incl. untuned and >80 shapes)."""
    rng = np.random.default_rng(sum(sizes))
    nr, nc, nk = 40, 36, 44
    ms, ns, ks = (workload.block_sizes(n, sizes, rng) for n in (nr, nc, nk))
    A = workload.random_panel(ms, ks, 0.3, rng)
    B = workload.random_panel(ks, ns, 0.3, rng)
    dm = DeviceMultiply(acc, ms, ns, ks, A.data.size, B.data.size, B.nblks, nthreads=2,
                        cfg=host.default_cfg(mm_stack_size=400, n_stacks=n_stacks), mode=host.LAUNCH | host.RECORD)
    dm.upload_panels(A.data, B.data, B.list3())
    dm.multiply(A.list3(), B.list3())
    prod = dm.download_c()
    assert any(not st["defined_mnk"] for st in dm.engine.stacks())  # the inhomogeneous path was really exercised
    dm.close()
    check_against_oracle(A, B, prod, ms, ns)


def test_first_touch_order_matches_index_oracle(acc):
    """One thread: the pre-finalize C index (order of first touch) equals the restated reference traversal."""
    rng = np.random.default_rng(3)
    ms = ns = ks = workload.block_sizes(64, [23], rng)
    A = workload.random_panel(ms, ks, 0.2, rng)
    B = workload.random_panel(ks, ns, 0.2, rng)
    prod, _ = multiply(acc, A, B, ms, ns, ks, nthreads=1, cfg=host.default_cfg(mm_stack_size=1000))
    ora = io.LocalMultiplyOracle(ms, ns, ks, mm_stack_size=1000)
    ora.multiply([tuple(int(v) for v in r) for r in A.list3()], [tuple(int(v) for v in r) for r in B.list3()])
    rows, cols, blk_p, _ = prod.parts[0]
    assert list(rows) == ora.c_row_i and list(cols) == ora.c_col_i and list(blk_p) == ora.c_blk_p
    check_against_oracle(A, B, prod, ms, ns)


def test_row_chunks_with_early_d2h(acc):
    """Engine with several row chunks per thread and pooled host result buffers: every chunk's D2H is enqueued behind its last
    stack (the bench's end-to-end path); results identical to the oracle, three multiplies in a row."""
    rng = np.random.default_rng(21)
    ms = ns = ks = workload.block_sizes(60, [13, 23], rng)
    A = workload.random_panel(ms, ks, 0.25, rng)
    B = workload.random_panel(ks, ns, 0.25, rng)
    dm = DeviceMultiply(acc, ms, ns, ks, A.data.size, B.data.size, B.nblks, nthreads=3, cfg=host.default_cfg(mm_stack_size=300, row_chunks=4))
    dm.upload_panels(A.data, B.data, B.list3())
    dm.multiply(A.list3(), B.list3())
    check_against_oracle(A, B, dm.download_c(), ms, ns)
    bufs = [np.empty(max(dm.engine.c_capacity(t), 1)) for t in range(3)]
    dm.set_result_buffers(bufs)
    for _ in range(2):
        dm.upload_panels(A.data, B.data, B.list3())
        dm.multiply(A.list3(), B.list3())
        check_against_oracle(A, B, dm.download_c(), ms, ns)
    dm.close()


def test_repeated_multiplies_on_pooled_buffers(acc):
    rng = np.random.default_rng(8)
    ms = ns = ks = workload.block_sizes(40, [13, 23], rng)
    A = workload.random_panel(ms, ks, 0.25, rng)
    B = workload.random_panel(ks, ns, 0.25, rng)
    dm = DeviceMultiply(acc, ms, ns, ks, A.data.size, B.data.size, B.nblks, nthreads=2, cfg=host.default_cfg(mm_stack_size=300))
    for _ in range(3):
        dm.upload_panels(A.data, B.data, B.list3())
        dm.multiply(A.list3(), B.list3())
        check_against_oracle(A, B, dm.download_c(), ms, ns)
    dm.close()


@pytest.mark.parametrize("cfgname,nblk", [("cfg2", 1000), ("cfg3", 400)])
def test_full_size_properties(acc, cfgname, nblk):
    """BASELINE.json sizes (cfg2: N=1000, 23x23, 10 %): sum(C) == sum_products colsum(A).rowsum(B) and the C block structure
    == boolean pattern product; also the recorded stacks drain to the same sum (stack-kernel-only path)."""
    w = workload.make_config(cfgname, nblk=nblk)
    A, B, bs = w["A"], w["B"], w["m_sizes"]
    n_st = 3 if cfgname == "cfg2" else 5
    dm = DeviceMultiply(acc, bs, bs, bs, A.data.size, B.data.size, B.nblks, nthreads=4, cfg=host.default_cfg(n_stacks=n_st),
                        mode=host.LAUNCH)
    dm.upload_panels(A.data, B.data, B.list3())
    dm.multiply(A.list3(), B.list3())
    prod = dm.download_c()
    flop = dm.engine.flop()
    dm.close()
    # expected number of products / flops from the patterns alone
    nb = bs.size
    amask = np.zeros((nb, nb), dtype=np.float32)
    amask[A.rows - 1, A.cols - 1] = 1
    bmask = np.zeros((nb, nb), dtype=np.float32)
    bmask[B.rows - 1, B.cols - 1] = 1
    cnt = amask @ bmask
    assert prod.nblks == int((cnt > 0).sum())
    w3 = (amask * bs[:, None].astype(np.float32) * bs[None, :]).astype(np.float64) @ (bmask * bs[None, :]).astype(np.float64)
    assert flop == int(round(2 * w3.sum()))
    # sum over all elements of C == sum over block pairs of colsum(A_ik) . rowsum(B_kj)
    colsum_a = [A.block(i).sum(axis=0) for i in range(A.nblks)]  # length k each
    rowsum_b = {}
    for i in range(B.nblks):
        rowsum_b.setdefault(int(B.rows[i]), []).append(B.block(i).sum(axis=1))
    rb = {k: np.sum(v, axis=0) for k, v in rowsum_b.items()}  # sum over all B blocks in block row k
    expected = sum(float(colsum_a[i] @ rb[int(A.cols[i])]) for i in range(A.nblks) if int(A.cols[i]) in rb)
    got = sum(float(p[3].sum()) for p in prod.parts)
    assert abs(got / expected - 1.0) <= 1e-10, (got, expected)


@pytest.mark.parametrize("fused", [False, True], ids=["norms_separate", "norms_fused_with_transpose"])
@pytest.mark.parametrize("nthreads", [1, 3])
def test_on_the_fly_filter(acc, nthreads, fused):
    """dbcsr_multiply(filter_eps=...): block norms from c_calculate_norms on the device (src/mm/dbcsr_mm_common.F:498-591),
    thresholds row_max_epss (src/mm/dbcsr_mm_cannon.F:1098-1107), products with a_norm*b_norm < row_eps skipped
    (src/mm/dbcsr_mm_csr.F:270-278).  Expected C = sum over exactly the surviving products, computed on the CPU."""
    rng = np.random.default_rng(77)
    sizes = [5, 13, 23]
    ms, ns, ks = (workload.block_sizes(n, sizes, rng) for n in (40, 36, 44))
    A = workload.random_panel(ms, ks, 0.35, rng)
    B = workload.random_panel(ks, ns, 0.35, rng)
    for P in (A, B):  # block magnitudes over four decades so that a mid-range eps cuts a good part of the products
        for i in range(P.nblks):
            P.block(i)[...] *= 10.0 ** rng.uniform(-4, 0)
    eps = 0.5
    dm = DeviceMultiply(acc, ms, ns, ks, A.data.size, B.data.size, B.nblks, nthreads=nthreads, cfg=host.default_cfg(mm_stack_size=400))
    try:
        # fused: the right panel's norms come out of the SAME pass that transposes it (libsmm_acc_b200_transpose_norms)
        dm.upload_panels(A.data, B.data, B.list3(), want_b_norms=fused)
        assert (dm.b_norms_fused is not None) == fused
        dm.multiply(A.list3(), B.list3(), filter_eps=eps)
        prod = dm.download_c()
        a_n, b_n = dm.a_norms, dm.b_norms
        flop = dm.engine.flop()
    finally:
        dm.close()
    # device norms = the oracle's (squared Frobenius norm, single precision), up to summation order
    ref_an = orc.norms(A.data, A.offsets, ms[A.rows - 1] * ks[A.cols - 1])
    ref_bn = orc.norms(B.data, B.offsets, ks[B.rows - 1] * ns[B.cols - 1])
    assert np.allclose(a_n, ref_an, rtol=1e-5, atol=0) and np.allclose(b_n, ref_bn, rtol=1e-5, atol=0)
    row_eps = io.row_max_epss(eps, np.bincount(A.rows - 1, minlength=ms.size))
    by_k = {}
    for j in range(B.nblks):
        by_k.setdefault(int(B.rows[j]), []).append(j)
    exp, kept, total, exp_flop = {}, 0, 0, 0
    for i in range(A.nblks):
        r = int(A.rows[i])
        for j in by_k.get(int(A.cols[i]), []):
            total += 1
            if np.float32(a_n[i] * b_n[j]) < row_eps[r - 1]:
                continue
            kept += 1
            key = (r, int(B.cols[j]))
            blk = A.block(i) @ B.block(j)
            exp[key] = exp[key] + blk if key in exp else blk
            exp_flop += 2 * blk.shape[0] * blk.shape[1] * A.block(i).shape[1]
    assert 0 < kept < total and flop == exp_flop
    got = prod.blocks()
    assert set(got.keys()) == set(exp.keys())
    num = sum(float(((got[k] - exp[k]) ** 2).sum()) for k in exp)
    den = sum(float((exp[k] ** 2).sum()) for k in exp)
    assert np.sqrt(num / den) <= 1e-10


def _preset_case(seed):
    rng = np.random.default_rng(seed)
    sizes = [5, 13, 23]
    ms, ns, ks = (workload.block_sizes(n, sizes, rng) for n in (40, 36, 44))
    A = workload.random_panel(ms, ks, 0.25, rng)
    B = workload.random_panel(ks, ns, 0.25, rng)
    C0 = workload.random_panel(ms, ns, 0.3, rng)
    return ms, ns, ks, A, B, C0


@pytest.mark.parametrize("nthreads", [1, 3])
def test_beta_flow_accumulates_onto_existing_c(acc, nthreads):
    """C = A*B + beta*C_old (dbcsr_multiply with beta != 0): the existing blocks are uploaded into the device work area and the
    stack kernels accumulate onto them (reference: zeroed device buffer + host block_add, src/mm/dbcsr_mm_accdrv.F:340-362)."""
    ms, ns, ks, A, B, C0 = _preset_case(5)
    beta = -0.75
    dm = DeviceMultiply(acc, ms, ns, ks, A.data.size, B.data.size, B.nblks, nthreads=nthreads, cfg=host.default_cfg(mm_stack_size=400))
    try:
        dm.upload_panels(A.data, B.data, B.list3())
        dm.multiply(A.list3(), B.list3(), c_preset=(C0.rows, C0.cols, beta * C0.data))
        prod = dm.download_c()
    finally:
        dm.close()
    C_in = orc.BlockMatrix(ms, ns, C0.rows, C0.cols, data=beta * C0.data)
    Cref = orc.multiply_blocks(to_oracle(A), to_oracle(B), C_in=C_in)
    got = prod.blocks()
    assert set(got.keys()) == set(zip(Cref.rows.tolist(), Cref.cols.tolist()))
    num = den = 0.0
    for (r, c, o) in zip(Cref.rows, Cref.cols, Cref.offsets):
        m, n = int(ms[r - 1]), int(ns[c - 1])
        ref = Cref.data[o:o + m * n].reshape(n, m).T
        num += float(((got[(int(r), int(c))] - ref) ** 2).sum())
        den += float((ref ** 2).sum())
    assert np.sqrt(num / den) <= 1e-10


def test_retain_sparsity(acc):
    """retain_sparsity: only the listed C blocks exist afterwards; each equals beta*C_old + sum of its products."""
    ms, ns, ks, A, B, C0 = _preset_case(6)
    dm = DeviceMultiply(acc, ms, ns, ks, A.data.size, B.data.size, B.nblks, nthreads=2, cfg=host.default_cfg(mm_stack_size=400))
    try:
        dm.upload_panels(A.data, B.data, B.list3())
        dm.multiply(A.list3(), B.list3(), c_preset=(C0.rows, C0.cols, C0.data), retain_sparsity=True)
        prod = dm.download_c()
    finally:
        dm.close()
    Cfull = orc.multiply_blocks(to_oracle(A), to_oracle(B), C_in=orc.BlockMatrix(ms, ns, C0.rows, C0.cols, data=C0.data))
    full = {(int(r), int(c)): Cfull.data[o:o + int(ms[r - 1]) * int(ns[c - 1])].reshape(int(ns[c - 1]), int(ms[r - 1])).T
            for r, c, o in zip(Cfull.rows, Cfull.cols, Cfull.offsets)}
    got = prod.blocks()
    listed = set(zip(C0.rows.tolist(), C0.cols.tolist()))
    assert set(got.keys()) == listed and len(listed) < len(full)
    num = sum(float(((got[k] - full[k]) ** 2).sum()) for k in listed)
    den = sum(float((full[k] ** 2).sum()) for k in listed)
    assert np.sqrt(num / den) <= 1e-10


@pytest.mark.parametrize("nthreads", [1, 3])
def test_final_filter_on_device(acc, nthreads):
    """multrec_filtering on the device before the download: same surviving block set as the restated reference filter applied
    to the unfiltered product, data of the survivors untouched and packed contiguously in index order."""
    rng = np.random.default_rng(9)
    sizes = [5, 13, 23]
    ms, ns, ks = (workload.block_sizes(n, sizes, rng) for n in (40, 36, 44))
    A = workload.random_panel(ms, ks, 0.2, rng)
    B = workload.random_panel(ks, ns, 0.2, rng)
    for i in range(A.nblks):  # row-dependent magnitudes => C block norms spread over many decades
        A.block(i)[...] *= 10.0 ** (-6.0 * (A.rows[i] % 7) / 6.0)
    eps = 1e-2
    dm = DeviceMultiply(acc, ms, ns, ks, A.data.size, B.data.size, B.nblks, nthreads=nthreads, cfg=host.default_cfg(mm_stack_size=400))
    try:
        dm.upload_panels(A.data, B.data, B.list3())
        dm.multiply(A.list3(), B.list3())
        full = dm.download_c()
        dm.filter_c(eps)
        filt = dm.download_c()
        d2h_full, d2h_filt = sum(8 * p[3].size for p in full.parts), dm.d2h_bytes
    finally:
        dm.close()
    assert len(filt.parts) == len(full.parts) == nthreads
    kept_total = 0
    for (rows, cols, blk_p, data), (frows, fcols, fblk_p, fdata) in zip(full.parts, filt.parts):
        er, ec, ep, enze, norms = io.multrec_filtering(eps, rows, cols, blk_p, ms, ns, data)
        # blocks whose norm sits within rounding of the threshold may legitimately differ (device summation order)
        assert all(abs(n2 - eps * eps) > 1e-9 * eps * eps for n2 in norms)
        assert frows.tolist() == er and fcols.tolist() == ec and fdata.size == enze
        nel = ms[frows - 1] * ns[fcols - 1]
        assert np.array_equal(fblk_p, 1 + np.concatenate([[0], np.cumsum(nel)[:-1]]).astype(np.int64)) if frows.size else True
        for j in range(frows.size):
            assert np.array_equal(fdata[fblk_p[j] - 1:fblk_p[j] - 1 + nel[j]], data[ep[j] - 1:ep[j] - 1 + nel[j]])
        kept_total += frows.size
    assert 0 < kept_total < full.nblks and d2h_filt < d2h_full


@pytest.mark.parametrize("nthreads,row_chunks", [(1, 1), (3, 2), (4, 4), (7, 3)])
def test_pipelined_left_panel_upload(acc, nthreads, row_chunks):
    """Left panel uploaded in the engine's block-row chunks with one event per chunk (stacks of a chunk wait only for their own
    rows): same product as the one-shot upload, including chunks without any A block and more chunks than block rows."""
    rng = np.random.default_rng(123 + nthreads)
    sizes = [5, 13, 23]
    ms, ns, ks = (workload.block_sizes(n, sizes, rng) for n in (20 if nthreads == 7 else 48, 40, 56))
    A = workload.random_panel(ms, ks, 0.3, rng)
    B = workload.random_panel(ks, ns, 0.3, rng)
    cfg = host.default_cfg(mm_stack_size=500, row_chunks=row_chunks)
    prod, flop = multiply(acc, A, B, ms, ns, ks, nthreads=nthreads, cfg=cfg, pipelined=True)
    check_against_oracle(A, B, prod, ms, ns)


def test_scheduler_host_driver_route(acc):
    """dbcsr_mm_sched_process: a stack the accelerator refuses (here: the inhomogeneous default stack with the reference's
    behaviour switched on, libsmm_acc_process returns -1 and leaves C untouched) goes to the HOST driver, whose contributions live
    in the host work matrix and are added to the downloaded device buffer at finalize (src/mm/dbcsr_mm_sched.F:340-363,
    src/mm/dbcsr_mm_accdrv.F:340-362).  The library has no CPU path: the driver is the caller's (here: the oracle's
    blas_process_mm_stack_d restatement on the host copies of the panels, B untransposed).  Without a driver the multiply fails."""
    from dbcsr_b200 import lib as acclib
    from dbcsr_b200.multiply import ProductC

    sizes = [5, 13, 23, 26, 32]
    rng = np.random.default_rng(77)
    nr, nc, nk = 40, 36, 44
    ms, ns, ks = (workload.block_sizes(n, sizes, rng) for n in (nr, nc, nk))
    A = workload.random_panel(ms, ks, 0.3, rng)
    B = workload.random_panel(ks, ns, 0.3, rng)
    nthreads = 2
    saved = acc.get_tunable("inhomogeneous")
    acc.set_tunable("inhomogeneous", 0)
    try:
        dm = DeviceMultiply(acc, ms, ns, ks, A.data.size, B.data.size, B.nblks, nthreads=nthreads, cfg=host.default_cfg(mm_stack_size=400, n_stacks=3))
        dm.upload_panels(A.data, B.data, B.list3())
        with pytest.raises(acclib.AccError):  # no host driver installed: the refusal is reported, nothing is computed on the CPU
            dm.multiply(A.list3(), B.list3())
        dm.engine.sync()
        host_c = [np.zeros(0) for _ in range(nthreads)]
        calls = []

        def driver(thread, m, n, k, defined_mnk, params7, c_datasize):
            if host_c[thread].size < c_datasize:
                host_c[thread] = np.concatenate([host_c[thread], np.zeros(c_datasize - host_c[thread].size)])
            orc.host_stack(params7, A.data, B.data, host_c[thread])  # in place
            calls.append((thread, defined_mnk, params7.shape[0]))
            return 0

        dm.engine.set_host_driver(driver)
        dm.upload_panels(A.data, B.data, B.list3())
        dm.multiply(A.list3(), B.list3())
        dev = dm.download_c()
        cpu = dm.engine.stats_cpu()
        assert calls and all(d == 0 for _, d, _ in calls) and cpu["stacks"] == len(calls) and cpu["entries"] == sum(c[2] for c in calls)
        rows_acc, tot_acc = dm.engine.stats()
        assert tot_acc["entries"] > 0  # the homogeneous stacks still ran on the accelerator
        prod = ProductC(ms, ns)
        for t, (rows, cols, blk_p, data) in enumerate(dev.parts):
            d = np.array(data, dtype=np.float64, copy=True)
            d[:min(d.size, host_c[t].size)] += host_c[t][:d.size]   # block_add of the two work areas
            prod.add(rows, cols, blk_p, d)
        dm.close()
    finally:
        acc.set_tunable("inhomogeneous", saved)
    check_against_oracle(A, B, prod, ms, ns)


@pytest.mark.parametrize("nthreads", [1, 3])
def test_device_c_buffer_grows_on_demand(acc, nthreads):
    """A device C buffer that starts far too small (c_capacity = 1000 elements) is enlarged while the stacks stream
    (dbcsr_data_ensure_size in the accelerator driver, src/mm/dbcsr_mm_accdrv.F:471-473): old contents kept, new tail zeroed;
    also across repeated multiplies on the pooled (now larger) buffer."""
    rng = np.random.default_rng(5)
    ms = ns = ks = workload.block_sizes(50, [13, 23], rng)
    A = workload.random_panel(ms, ks, 0.3, rng)
    B = workload.random_panel(ks, ns, 0.3, rng)
    dm = DeviceMultiply(acc, ms, ns, ks, A.data.size, B.data.size, B.nblks, nthreads=nthreads, cfg=host.default_cfg(mm_stack_size=200), c_capacity=1000)
    for _ in range(2):
        dm.upload_panels(A.data, B.data, B.list3())
        dm.multiply(A.list3(), B.list3())
        check_against_oracle(A, B, dm.download_c(), ms, ns)
    assert all(dm.engine.c_capacity(t) >= dm.engine.c_index(t)[3] for t in range(nthreads))
    dm.close()
