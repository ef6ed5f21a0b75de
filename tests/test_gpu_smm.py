"""GPU parity tests of the stack-drain hot path, called through the C ABI (libsmm_acc_process & co) and checked against the
oracle (oracle/dbcsr_oracle.c).  Modelled on the reference's own unit tests:
  tests/libsmm_acc_unittest_multiply.cpp.template + src/acc/libsmm_acc/libsmm_acc_benchmark.cpp:45-52,103-170,287-291
  (integer-valued inputs => exact checksum comparison), tests/libsmm_acc_unittest_transpose.cpp.
FP64 tolerance for real-valued inputs: relative Frobenius error <= 1e-10 (BASELINE.json north_star); integer inputs: bit-exact.
"""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

TUNED = [5, 13, 23, 26, 32]


@pytest.fixture(scope="module")
def acc():
    from dbcsr_b200 import lib as acclib

    a = acclib.Acc(0)
    a.s = a.stream_create("test", 0)
    yield a
    a.stream_destroy(a.s)
    a.finalize()


def run_process(acc, stack3, a, b, c_size, m, n, k, host7=None, def_mnk=True, pad_elems=0):
    """Upload, run libsmm_acc_process on a device stack, download C. `pad_elems` doubles are put in front of A/B so that
    block offsets start at an odd element (8 mod 16 byte alignment)."""
    stack3 = np.ascontiguousarray(stack3, dtype=np.int32).reshape(-1, 3)
    if pad_elems:
        stack3 = stack3.copy()
        stack3[:, 0] += pad_elems
        stack3[:, 1] += pad_elems
        a = np.concatenate([np.full(pad_elems, np.nan), a])
        b = np.concatenate([np.full(pad_elems, np.nan), b])
    d_a, d_b = acc.to_device(a, acc.s), acc.to_device(b, acc.s)
    d_s = acc.to_device(stack3, acc.s)
    d_c = acc.dev_alloc(c_size * 8)
    acc.memset_zero(d_c, acc.s)
    if host7 is None:
        host7 = np.zeros((stack3.shape[0], 7), dtype=np.int32)
        host7[:, 0], host7[:, 1], host7[:, 2] = m, n, k
        host7[:, 3:6] = stack3
    rc = acc.process(host7, d_s.ptr, stack3.shape[0], d_a.ptr, d_b.ptr, d_c.ptr, m, n, k, def_mnk, acc.s, acc.s)
    c = acc.to_host(d_c, (c_size,), np.float64, acc.s)
    for d in (d_a, d_b, d_s, d_c):
        d.free()
    return rc, c


def ref_test_problem(m, n, k, n_a=100, n_b=100, n_c=10, n_stack=100, seed=1):
    """The reference's `test` benchmark problem (libsmm_acc_benchmark.cpp:45-52): matInit(42/24) + rand() stack."""
    L = orc.lib()
    a, b = np.empty(n_a * m * k), np.empty(n_b * k * n)
    L.orc_mat_init(a, n_a, m, k, 42)
    L.orc_mat_init(b, n_b, k, n, 24)
    stack = np.empty(3 * n_stack, dtype=np.int32)
    orc.srand(seed)
    L.orc_stack_init(stack, n_stack, n_c, n_a, n_b, m, n, k)
    return a, b, stack.reshape(-1, 3), n_c * m * n


@pytest.mark.parametrize("m", TUNED)
def test_tuned_triplets_exact_checksum(acc, m):
    """All 25 (n,k) per m: integer-valued inputs, exact equality of every C element and of checkSum (reference :287-291)."""
    for n in TUNED:
        for k in TUNED:
            a, b, stack, csz = ref_test_problem(m, n, k)
            rc, c = run_process(acc, stack, a, b, csz, m, n, k)
            assert rc == 0, (m, n, k, rc)
            c_ref = orc.stack_calc(stack, np.zeros(csz), a, b, m, n, k)
            assert np.array_equal(c, c_ref), (m, n, k, np.abs(c - c_ref).max())
            assert orc.lib().orc_checksum(c, csz // (m * n), m, n) == orc.lib().orc_checksum(c_ref, csz // (m * n), m, n)


@pytest.mark.parametrize("mnk", [(23, 23, 23), (5, 5, 5), (13, 13, 13), (26, 26, 26), (32, 32, 32), (23, 5, 32), (13, 32, 5), (32, 13, 26)])
def test_random_values_frobenius(acc, mnk):
    """Uniform(0,1) data (dlarnv-like), 16005-entry C-sorted stack of the reference's timing problem scaled down:
    relative Frobenius error <= 1e-10 (north_star tolerance; observed ~1e-16)."""
    m, n, k = mnk
    rng = np.random.default_rng(7)
    n_a, n_b, n_c, S = 2000, 2000, 300, 4000
    a, b = rng.random(n_a * m * k), rng.random(n_b * k * n)
    stack = np.empty(3 * S, dtype=np.int32)
    orc.srand(3)
    orc.lib().orc_stack_init(stack, S, n_c, n_a, n_b, m, n, k)
    stack = stack.reshape(-1, 3)
    for pad in (0, 1):  # pad=1: every block of an even-sized shape sits at 8 mod 16 bytes, odd-sized ones alternate
        rc, c = run_process(acc, stack, a, b, n_c * m * n, m, n, k, pad_elems=pad)
        assert rc == 0
        c_ref = orc.stack_calc(stack, np.zeros(n_c * m * n), a, b, m, n, k)
        err = np.linalg.norm(c - c_ref) / np.linalg.norm(c_ref)
        assert err <= 1e-10, (mnk, pad, err)


def test_last_block_of_allocation_and_alignment(acc):
    """A and B sized exactly to their blocks: the last block ends at the end of the allocation (the TMA window would over-read
    8 bytes there; the kernel must take the guarded path) and odd block sizes alternate between 0 and 8 mod 16 alignment."""
    m = n = k = 23
    rng = np.random.default_rng(11)
    for nblk in (1, 2, 3, 8):
        a, b = rng.random(nblk * m * k), rng.random(nblk * k * n)
        S = 4 * nblk
        stack = np.zeros((S, 3), dtype=np.int32)
        stack[:, 0] = (np.arange(S) % nblk) * m * k + 1
        stack[:, 1] = ((np.arange(S) // 2) % nblk) * k * n + 1
        stack[:, 2] = (np.arange(S) // 4) * m * n + 1
        # make sure the very last block of both panels is used
        stack[-1, 0] = (nblk - 1) * m * k + 1
        stack[-1, 1] = (nblk - 1) * k * n + 1
        rc, c = run_process(acc, stack, a, b, nblk * m * n, m, n, k)
        assert rc == 0
        c_ref = orc.stack_calc(stack, np.zeros(nblk * m * n), a, b, m, n, k)
        assert np.linalg.norm(c - c_ref) / np.linalg.norm(c_ref) <= 1e-13


def test_edge_stacks(acc):
    """Empty stack, single entry, unsorted stack (atomics make order irrelevant), one C block hit by every entry,
    and a full 30000-entry stack (MM_STACK_SIZE of GPU builds, src/core/dbcsr_config.F:76-82)."""
    m = n = k = 23
    rng = np.random.default_rng(5)
    n_a = n_b = 500
    a, b = rng.random(n_a * m * k), rng.random(n_b * k * n)
    # empty
    rc, c = run_process(acc, np.zeros((0, 3), dtype=np.int32), a, b, m * n, m, n, k)
    assert rc == 0 and not c.any()
    for S, n_c, shuffle in [(1, 1, False), (7, 7, False), (1000, 1, False), (5000, 50, True), (30000, 3000, False)]:
        stack = np.zeros((S, 3), dtype=np.int32)
        stack[:, 0] = rng.integers(0, n_a, S) * m * k + 1
        stack[:, 1] = rng.integers(0, n_b, S) * k * n + 1
        stack[:, 2] = np.sort(rng.integers(0, n_c, S)) * m * n + 1
        if shuffle:
            stack = stack[rng.permutation(S)]
        rc, c = run_process(acc, stack, a, b, n_c * m * n, m, n, k)
        assert rc == 0
        c_ref = orc.stack_calc(stack, np.zeros(n_c * m * n), a, b, m, n, k)
        assert np.linalg.norm(c - c_ref) / max(np.linalg.norm(c_ref), 1e-300) <= 1e-10, (S, n_c)


def test_accumulates_into_existing_c(acc):
    """C += A*B: a second process call on the same C buffer adds on top (device C buffer lives across stacks,
    src/mm/dbcsr_mm_accdrv.F:209-216)."""
    m, n, k = 13, 23, 5
    a, b, stack, csz = ref_test_problem(m, n, k)
    d_a, d_b, d_s = acc.to_device(a, acc.s), acc.to_device(b, acc.s), acc.to_device(stack, acc.s)
    d_c = acc.to_device(np.full(csz, 1.5), acc.s)
    for _ in range(3):
        assert acc.process(None, d_s.ptr, stack.shape[0], d_a.ptr, d_b.ptr, d_c.ptr, m, n, k, True, acc.s, acc.s) == 0
    c = acc.to_host(d_c, (csz,), np.float64, acc.s)
    c_ref = np.full(csz, 1.5)
    for _ in range(3):
        orc.stack_calc(stack, c_ref, a, b, m, n, k)
    assert np.array_equal(c, c_ref)


@pytest.mark.parametrize("mnk", [(7, 9, 11), (1, 1, 1), (4, 4, 4), (14, 29, 32), (45, 67, 78), (80, 80, 80), (3, 64, 2)])
def test_generic_kernel_shapes(acc, mnk):
    """Shapes without a specialised kernel run on the generic kernel and report 10 (= untuned, libsmm_acc.cpp:319)."""
    m, n, k = mnk
    a, b, stack, csz = ref_test_problem(m, n, k, n_a=20, n_b=20, n_c=5, n_stack=40)
    rc, c = run_process(acc, stack, a, b, csz, m, n, k)
    assert rc == 10
    c_ref = orc.stack_calc(stack, np.zeros(csz), a, b, m, n, k)
    assert np.array_equal(c, c_ref)


@pytest.mark.parametrize("hugedmma", [1, 0], ids=["panel_dmma_kernel", "scalar_generic_kernel"])
def test_large_blocks_b_not_transposed(acc, hugedmma):
    """Any dim > max_kernel_dim: B is transposed by DBCSR only when BOTH n and k fit max_kernel_dim (src/acc/libsmm_acc/libsmm_acc.cpp:267-270).
    Default: the panel DMMA kernel (smm_dmma_huge.cuh; C panels of <= 96 x 80, K chunks of 32, ragged edges); hugedmma = 0: the scalar kernel."""
    rng = np.random.default_rng(3)
    saved = acc.get_tunable("hugedmma")
    acc.set_tunable("hugedmma", hugedmma)
    shapes = [(100, 5, 90), (5, 100, 7), (81, 81, 81), (100, 100, 100), (200, 96, 161), (97, 81, 33), (120, 40, 70), (83, 7, 3), (1, 1, 130)]
    for (m, n, k) in shapes:
        n_a = n_b = 6
        n_c = 3
        S = 12
        a, b = rng.random(n_a * m * k), rng.random(n_b * k * n)
        host = np.zeros((S, 7), dtype=np.int32)
        host[:, 0], host[:, 1], host[:, 2] = m, n, k
        host[:, 3] = rng.integers(0, n_a, S) * m * k + 1
        host[:, 4] = rng.integers(0, n_b, S) * k * n + 1
        cb = np.sort(rng.integers(0, n_c, S))
        host[:, 5], host[:, 6] = cb * m * n + 1, cb + 1
        b_dev = b.copy()
        if n <= 80 and k <= 80:
            orc.transpose_blocks(np.arange(n_b, dtype=np.int32) * k * n, b_dev, k, n)
        rc, c = run_process(acc, host[:, 3:6].copy(), a, b_dev, n_c * m * n, m, n, k, host7=host)
        assert rc == 10
        c_ref = orc.host_stack(host, a, b, np.zeros(n_c * m * n))
        assert np.linalg.norm(c - c_ref) / np.linalg.norm(c_ref) <= 1e-12, (m, n, k)
    acc.set_tunable("hugedmma", saved)


def test_unsupported_requests_leave_c_untouched(acc):
    """Unknown dtype -> -10 (reference return code, libsmm_acc.cpp:338); def_mnk=0 without a host stack -> -1.  C must not be
    modified so that DBCSR can redo the stack on the CPU."""
    m = n = k = 23
    a, b, stack, csz = ref_test_problem(m, n, k)
    d_a0, d_b0, d_s0 = acc.to_device(a, acc.s), acc.to_device(b, acc.s), acc.to_device(stack, acc.s)
    d_c0 = acc.dev_alloc(csz * 8)
    acc.memset_zero(d_c0, acc.s)
    assert acc.process(None, d_s0.ptr, stack.shape[0], d_a0.ptr, d_b0.ptr, d_c0.ptr, m, n, k, False, acc.s, acc.s) == -1
    for dt in (0, 2, 11):
        assert acc.process(None, d_s0.ptr, stack.shape[0], d_a0.ptr, d_b0.ptr, d_c0.ptr, m, n, k, True, acc.s, acc.s, datatype=dt) == -10
    assert not acc.to_host(d_c0, (csz,), np.float64, acc.s).any()
    # with the host stack the homogeneous problem gives the same result through the inhomogeneous path
    rc, c = run_process(acc, stack, a, b, csz, m, n, k, def_mnk=False)
    assert rc == 0 and np.array_equal(c, orc.stack_calc(stack, np.zeros(csz), a, b, m, n, k))


@pytest.mark.parametrize("dt,npt,tol", [(1, np.float32, 2e-6), (5, np.complex64, 2e-6), (7, np.complex128, 1e-14)])
def test_other_abi_types(acc, dt, npt, tol):
    """real_4 / complex_4 / complex_8 (libsmm_acc_data_t 1/5/7): the reference answers -10 (CPU); here transpose + process run on the
    device with the typed generic kernel and report 10 (untuned).  Checked against numpy in double precision."""
    rng = np.random.default_rng(dt)
    for (m, n, k) in [(23, 23, 23), (5, 13, 7), (32, 4, 9)]:
        n_a = n_b = 30
        n_c, S = 6, 120
        def rnd(sz):
            x = rng.random(sz)
            return (x + 1j * rng.random(sz)).astype(npt) if np.issubdtype(npt, np.complexfloating) else x.astype(npt)
        a, b = rnd(n_a * m * k), rnd(n_b * k * n)              # b: k x n col-major blocks (host layout)
        stack = np.zeros((S, 3), dtype=np.int32)
        ia, ib = rng.integers(0, n_a, S), rng.integers(0, n_b, S)
        ic = np.sort(rng.integers(0, n_c, S))
        stack[:, 0], stack[:, 1], stack[:, 2] = ia * m * k + 1, ib * k * n + 1, ic * m * n + 1
        d_a, d_b, d_s = acc.to_device(a, acc.s), acc.to_device(b, acc.s), acc.to_device(stack, acc.s)
        d_t = acc.to_device(np.arange(n_b, dtype=np.int32) * k * n, acc.s)
        acc.transpose(d_t.ptr, 0, n_b, d_b.ptr, k, n, acc.s, datatype=dt)
        bt = acc.to_host(d_b, b.shape, npt, acc.s)
        assert np.array_equal(bt.reshape(n_b, k, n), b.reshape(n_b, n, k).transpose(0, 2, 1))  # n x k col-major now
        d_c = acc.dev_alloc(n_c * m * n * a.itemsize)
        acc.memset_zero(d_c, acc.s)
        assert acc.process(None, d_s.ptr, S, d_a.ptr, d_b.ptr, d_c.ptr, m, n, k, True, acc.s, acc.s, datatype=dt) == 10
        c = acc.to_host(d_c, (n_c * m * n,), npt, acc.s)
        hp = np.complex128 if np.issubdtype(npt, np.complexfloating) else np.float64
        c_ref = np.zeros((n_c, n, m), dtype=hp)  # col-major blocks: [blk][col][row]
        A = a.astype(hp).reshape(n_a, k, m)       # [blk][l][row]
        B = b.astype(hp).reshape(n_b, n, k)       # [blk][col][l]
        for e in range(S):
            c_ref[ic[e]] += B[ib[e]] @ A[ia[e]]
        err = np.linalg.norm(c.astype(hp) - c_ref.reshape(-1)) / np.linalg.norm(c_ref)
        assert err <= tol, (dt, m, n, k, err)
        for d in (d_a, d_b, d_s, d_t, d_c):
            d.free()


def test_transpose_all_tuned_pairs(acc):
    """libsmm_acc_transpose vs the oracle for every (m,n) pair incl. offset argument (tests/libsmm_acc_unittest_transpose.cpp)."""
    for m in TUNED + [1, 7, 80]:
        for n in TUNED + [1, 80]:
            nblk = 37
            mat = np.empty(nblk * m * n)
            orc.lib().orc_mat_init(mat, nblk, m, n, 42)
            offs = (np.arange(nblk, dtype=np.int32) * m * n).astype(np.int32)
            skip = 5
            d_m, d_o = acc.to_device(mat, acc.s), acc.to_device(offs, acc.s)
            acc.transpose(d_o.ptr, skip, nblk - skip, d_m.ptr, m, n, acc.s)
            out = acc.to_host(d_m, mat.shape, np.float64, acc.s)
            ref = mat.copy()
            orc.transpose_blocks(offs[skip:], ref, m, n)
            assert np.array_equal(out, ref), (m, n)
            d_m.free()
            d_o.free()


def test_transpose_then_process_matches_host_path(acc):
    """DBCSR flow: right panel uploaded untransposed, libsmm_acc_transpose in place, then process == CPU path on host stacks."""
    m, n, k = 23, 13, 26
    rng = np.random.default_rng(9)
    n_a = n_b = 64
    n_c, S = 16, 256
    a, b = rng.random(n_a * m * k), rng.random(n_b * k * n)
    host = np.zeros((S, 7), dtype=np.int32)
    host[:, 0], host[:, 1], host[:, 2] = m, n, k
    host[:, 3] = rng.integers(0, n_a, S) * m * k + 1
    host[:, 4] = rng.integers(0, n_b, S) * k * n + 1
    cb = rng.integers(0, n_c, S)
    host[:, 5], host[:, 6] = cb * m * n + 1, cb + 1
    dev3 = host[np.argsort(host[:, 5], kind="stable")][:, 3:6].copy()  # stack_sort, src/mm/dbcsr_mm_accdrv.F:364-384
    d_a, d_b, d_s = acc.to_device(a, acc.s), acc.to_device(b, acc.s), acc.to_device(dev3, acc.s)
    d_t = acc.to_device((np.arange(n_b, dtype=np.int32) * k * n), acc.s)
    acc.transpose(d_t.ptr, 0, n_b, d_b.ptr, k, n, acc.s)
    d_c = acc.dev_alloc(n_c * m * n * 8)
    acc.memset_zero(d_c, acc.s)
    assert acc.process(host, d_s.ptr, S, d_a.ptr, d_b.ptr, d_c.ptr, m, n, k, True, acc.s, acc.s) == 0
    c = acc.to_host(d_c, (n_c * m * n,), np.float64, acc.s)
    c_ref = orc.host_stack(host, a, b, np.zeros(n_c * m * n))
    assert np.linalg.norm(c - c_ref) / np.linalg.norm(c_ref) <= 1e-10


def test_norms(acc):
    rng = np.random.default_rng(2)
    nblk = 1000
    sizes = rng.choice([25, 169, 529, 676, 1024, 1], nblk).astype(np.int32)
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int32)
    mat = rng.random(int(sizes.sum()))
    d_m, d_o, d_n = acc.to_device(mat, acc.s), acc.to_device(offs, acc.s), acc.to_device(sizes, acc.s)
    d_out = acc.dev_alloc(nblk * 4)
    acc.norms(d_m.ptr, nblk, d_o.ptr, d_n.ptr, d_out.ptr, acc.s)
    out = acc.to_host(d_out, (nblk,), np.float32, acc.s)
    ref = orc.norms(mat, offs, sizes)
    assert np.allclose(out, ref, rtol=2e-7, atol=0)


def test_concurrent_threads_own_streams(acc):
    """DBCSR calls libsmm_acc_process from every OpenMP thread on its own stream and C buffer (src/mm/dbcsr_mm_accdrv.F:229,300);
    libsmm_acc_is_thread_safe() promises that works.  8 host threads x 20 stacks each, shared A/B, private C."""
    import threading

    m = n = k = 23
    rng = np.random.default_rng(17)
    n_a = n_b = 400
    a, b = rng.random(n_a * m * k), rng.random(n_b * k * n)
    d_a, d_b = acc.to_device(a, acc.s), acc.to_device(b, acc.s)
    nthreads, nstacks, S, n_c = 8, 20, 3000, 60
    results, errors = [None] * nthreads, []

    def worker(t):
        try:
            r = np.random.default_rng(100 + t)
            s = acc.stream_create("t%d" % t, 0)
            d_c = acc.dev_alloc(n_c * m * n * 8)
            acc.memset_zero(d_c, s)
            c_ref = np.zeros(n_c * m * n)
            keep = []
            for _ in range(nstacks):
                st = np.zeros((S, 3), dtype=np.int32)
                st[:, 0] = r.integers(0, n_a, S) * m * k + 1
                st[:, 1] = r.integers(0, n_b, S) * k * n + 1
                st[:, 2] = np.sort(r.integers(0, n_c, S)) * m * n + 1
                d_s = acc.dev_alloc(st.nbytes)
                keep.append((d_s, acc.h2d(st, d_s, s)))
                rc = acc.process(None, d_s.ptr, S, d_a.ptr, d_b.ptr, d_c.ptr, m, n, k, True, s, s)
                assert rc == 0
                orc.stack_calc(st, c_ref, a, b, m, n, k)
            c = acc.to_host(d_c, (n_c * m * n,), np.float64, s)
            results[t] = np.linalg.norm(c - c_ref) / np.linalg.norm(c_ref)
            for d_s, _ in keep:
                d_s.free()
            d_c.free()
            acc.stream_destroy(s)
        except Exception as ex:  # pragma: no cover
            errors.append(repr(ex))

    ths = [threading.Thread(target=worker, args=(t,)) for t in range(nthreads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errors, errors
    assert all(r is not None and r <= 1e-10 for r in results), results


@pytest.mark.parametrize("mnk", [(23, 23, 23), (13, 13, 13), (23, 13, 5)])
def test_launch_knobs_keep_results(acc, mnk):
    """Run-time launch knobs (chunk size, run-aligned chunk boundaries, equal chunks; include/dbcsr_acc_libsmm.h) only change
    how a stack is split over the warps -- and, for 23^3, the flush of a run goes through one TMA bulk reduction
    (cp.reduce.async.bulk.add.f64) instead of per-element REDs: every combination must give the oracle's result exactly on
    integer-valued inputs, for C-sorted stacks with short and long runs, unsorted stacks and stacks smaller than the grid."""
    m, n, k = mnk
    rng = np.random.default_rng(17)
    n_a = n_b = 300
    a = rng.integers(0, 4, n_a * m * k).astype(np.float64)
    b = rng.integers(0, 4, n_b * k * n).astype(np.float64)
    saved = {name: acc.get_tunable(name) for name in ("balance", "align", "chunk")}
    try:
        for S, n_c, shuffle in [(3, 2, False), (200, 120, False), (4000, 37, False), (30000, 17000, False), (9000, 300, True)]:
            stack = np.zeros((S, 3), dtype=np.int32)
            stack[:, 0] = rng.integers(0, n_a, S) * m * k + 1
            stack[:, 1] = rng.integers(0, n_b, S) * k * n + 1
            stack[:, 2] = np.sort(rng.integers(0, n_c, S)) * m * n + 1
            if shuffle:
                stack = stack[rng.permutation(S)]
            c_ref = orc.stack_calc(stack, np.zeros(n_c * m * n), a, b, m, n, k)
            for balance, align, chunk in [(0, -1, -1), (0, 0, 0), (0, 1, 0), (1, 1, 0), (0, 1, 5), (1, 0, 12), (0, 1, 40)]:
                acc.set_tunable("balance", balance)
                acc.set_tunable("align", align)
                acc.set_tunable("chunk", chunk)
                rc, c = run_process(acc, stack, a, b, n_c * m * n, m, n, k, pad_elems=1 if chunk == 5 else 0)
                assert rc == 0
                assert np.array_equal(c, c_ref), (mnk, S, n_c, shuffle, balance, align, chunk, float(np.abs(c - c_ref).max()))
    finally:
        for name, v in saved.items():
            acc.set_tunable(name, v)


@pytest.mark.parametrize("chain", [False, True])
def test_producer_kernels_in_front_of_process_on_one_stream(acc, chain):
    """Every stack kernel is launched with programmatic stream serialization.  Whatever precedes it on the SAME stream -- a memset
    of C, the in-place transpose of the right panel, an earlier drain -- must be complete and visible before the kernel reads:
    no host synchronisation anywhere in the loop, 40 rounds of memset -> transpose -> process -> process -> transpose back, on
    integer-valued data (exact).  chain=True declares the stream a chain of independent drains (libsmm_acc_b200_stream_chain):
    the library itself falls back to the waiting mode behind its own transpose kernels."""
    m = n = k = 23
    rng = np.random.default_rng(31)
    n_a = n_b = 400
    n_c, S = 1500, 20000
    a = rng.integers(0, 4, n_a * m * k).astype(np.float64)
    b = rng.integers(0, 4, n_b * k * n).astype(np.float64)   # untransposed k x n blocks
    stack = np.zeros((S, 3), dtype=np.int32)
    stack[:, 0] = rng.integers(0, n_a, S) * m * k + 1
    stack[:, 1] = rng.integers(0, n_b, S) * k * n + 1
    stack[:, 2] = np.sort(rng.integers(0, n_c, S)) * m * n + 1
    host7 = np.zeros((S, 7), dtype=np.int32)
    host7[:, 0], host7[:, 1], host7[:, 2] = m, n, k
    host7[:, 3:6] = stack
    c_ref = 2.0 * orc.host_stack(host7, a, b, np.zeros(n_c * m * n))  # two drains per round
    d_a, d_b, d_s = acc.to_device(a, acc.s), acc.to_device(b, acc.s), acc.to_device(stack, acc.s)
    d_t = acc.to_device((np.arange(n_b, dtype=np.int32) * k * n), acc.s)
    d_c = acc.dev_alloc(n_c * m * n * 8)
    out = acc.host_alloc((n_c * m * n,), np.float64)
    acc.stream_chain(acc.s, chain)
    try:
        for rnd in range(40):
            acc.memset_zero(d_c, acc.s)
            acc.transpose(d_t.ptr, 0, n_b, d_b.ptr, k, n, acc.s)           # B -> Bt in place
            for _ in range(2):
                assert acc.process(None, d_s.ptr, S, d_a.ptr, d_b.ptr, d_c.ptr, m, n, k, True, acc.s, acc.s) == 0
            acc.transpose(d_t.ptr, 0, n_b, d_b.ptr, n, k, acc.s)           # back to k x n for the next round
            if rnd % 8 == 7:
                acc.d2h(d_c, out.array, acc.s)
                acc.stream_sync(acc.s)
                assert np.array_equal(out.array, c_ref), (rnd, float(np.abs(out.array - c_ref).max()))
    finally:
        acc.stream_chain(acc.s, False)
    for d in (d_a, d_b, d_s, d_t, d_c):
        d.free()
    out.free()


def test_chain_mode_many_drains_exact(acc):
    """Chain mode proper: 60 drains of different stacks back to back into one C buffer, nothing else on the stream."""
    m = n = k = 23
    rng = np.random.default_rng(32)
    n_a = n_b = 500
    n_c, S, reps = 4000, 30000, 60
    a = rng.integers(0, 3, n_a * m * k).astype(np.float64)
    bt = rng.integers(0, 3, n_b * k * n).astype(np.float64)
    stacks = []
    c_ref = np.zeros(n_c * m * n)
    for r in range(3):
        st = np.zeros((S, 3), dtype=np.int32)
        st[:, 0] = rng.integers(0, n_a, S) * m * k + 1
        st[:, 1] = rng.integers(0, n_b, S) * k * n + 1
        st[:, 2] = np.sort(rng.integers(0, n_c, S)) * m * n + 1
        stacks.append(st)
        c_ref += (reps // 3) * orc.stack_calc(st, np.zeros(n_c * m * n), a, bt, m, n, k)
    d_a, d_b = acc.to_device(a, acc.s), acc.to_device(bt, acc.s)
    d_s = [acc.to_device(st, acc.s) for st in stacks]
    d_c = acc.dev_alloc(n_c * m * n * 8)
    acc.memset_zero(d_c, acc.s)
    acc.stream_chain(acc.s, True)
    try:
        for r in range(reps):
            assert acc.process(None, d_s[r % 3].ptr, S, d_a.ptr, d_b.ptr, d_c.ptr, m, n, k, True, acc.s, acc.s) == 0
        c = acc.to_host(d_c, (n_c * m * n,), np.float64, acc.s)
    finally:
        acc.stream_chain(acc.s, False)
    assert np.array_equal(c, c_ref), float(np.abs(c - c_ref).max())
    for d in [d_a, d_b, d_c] + d_s:
        d.free()


@pytest.mark.parametrize("mnk", [(5, 5, 5), (5, 5, 32), (5, 13, 23), (13, 5, 26), (5, 13, 5), (13, 5, 13)])
def test_tiny_block_shapes_long_and_short_runs(acc, mnk):
    """Shapes with m*n <= 96 (candidates of the lane-per-element kernel, smm_tiny.cuh; which kernel runs is the autotune
    database's choice): exact on integer data for sorted stacks with long runs, short runs, unsorted stacks and tiny stacks."""
    m, n, k = mnk
    rng = np.random.default_rng(33)
    n_a = n_b = 700
    a = rng.integers(0, 4, n_a * m * k).astype(np.float64)
    bt = rng.integers(0, 4, n_b * k * n).astype(np.float64)
    for S, n_c, shuffle in [(1, 1, False), (33, 2, False), (5000, 4000, False), (30000, 9, False), (30000, 30000, False), (7000, 500, True)]:
        stack = np.zeros((S, 3), dtype=np.int32)
        stack[:, 0] = rng.integers(0, n_a, S) * m * k + 1
        stack[:, 1] = rng.integers(0, n_b, S) * k * n + 1
        stack[:, 2] = np.sort(rng.integers(0, n_c, S)) * m * n + 1
        if shuffle:
            stack = stack[rng.permutation(S)]
        rc, c = run_process(acc, stack, a, bt, n_c * m * n, m, n, k, pad_elems=1 if S == 5000 else 0)
        assert rc == 0
        c_ref = orc.stack_calc(stack, np.zeros(n_c * m * n), a, bt, m, n, k)
        assert np.array_equal(c, c_ref), (mnk, S, n_c, shuffle, float(np.abs(c - c_ref).max()))


def test_transpose_norms_fused_matches_separate_passes(acc):
    """libsmm_acc_b200_transpose_norms = libsmm_acc_transpose + c_calculate_norms in one pass: the transposed panel is identical,
    the norms land at the list positions given by dev_trs_blk and equal the oracle's (float sum of squares); offset argument;
    blocks above max_kernel_dim are refused with -3 and left untouched."""
    rng = np.random.default_rng(41)
    for m, n in [(23, 23), (5, 13), (32, 7), (80, 80), (1, 1)]:
        nblk, skip = 53, 4
        mat = rng.random(nblk * m * n) * 10.0 ** rng.uniform(-3, 3, nblk * m * n)
        offs = (np.arange(nblk, dtype=np.int32) * m * n).astype(np.int32)
        perm = rng.permutation(nblk).astype(np.int32)  # transpose-stack position -> list position
        d_m, d_o, d_p = acc.to_device(mat, acc.s), acc.to_device(offs, acc.s), acc.to_device(perm, acc.s)
        d_n = acc.dev_alloc(4 * nblk)
        acc.h2d(np.full(nblk, np.float32(-1.0)), d_n, acc.s)
        rc = acc.L.libsmm_acc_b200_transpose_norms(d_o.ptr, d_p.ptr, skip, nblk - skip, d_m.ptr, m, n, 80, d_n.ptr, acc.s)
        assert rc == 0
        out = acc.to_host(d_m, mat.shape, np.float64, acc.s)
        norms = acc.to_host(d_n, (nblk,), np.float32, acc.s)
        ref = mat.copy()
        orc.transpose_blocks(offs[skip:], ref, m, n)
        assert np.array_equal(out, ref), (m, n)
        ref_n = orc.norms(mat, offs, np.full(nblk, m * n, dtype=np.int32))
        touched = np.zeros(nblk, dtype=bool)
        touched[perm[skip:]] = True
        assert np.allclose(norms[perm[skip:]], ref_n[skip:], rtol=1e-5, atol=0), (m, n)
        assert np.all(norms[~touched] == -1.0)
        for d in (d_m, d_o, d_p, d_n):
            d.free()
    d_m = acc.to_device(np.ones(81 * 81), acc.s)
    d_o = acc.to_device(np.zeros(1, dtype=np.int32), acc.s)
    d_n = acc.dev_alloc(4)
    assert acc.L.libsmm_acc_b200_transpose_norms(d_o.ptr, None, 0, 1, d_m.ptr, 81, 81, 80, d_n.ptr, acc.s) == -3
    for d in (d_m, d_o, d_n):
        d.free()


def test_memset_zero_trickle(acc):
    """Bounded-rate zeroing (libsmm_acc_b200_memset_zero_trickle): same result as c_dbcsr_acc_memset_zero, any CTA count, offset."""
    n = 1 << 20
    d = acc.dev_alloc(8 * n)
    for nctas, off in [(1, 0), (8, 16 * 1000), (64, 0)]:
        acc.h2d(np.full(n, 3.25), d, acc.s)
        rc = acc.L.libsmm_acc_b200_memset_zero_trickle(d.ptr, off, 8 * n - off - 32, nctas, acc.s)
        assert rc == 0
        out = acc.to_host(d, (n,), np.float64, acc.s)
        exp = np.full(n, 3.25)
        exp[off // 8:n - 4] = 0.0
        assert np.array_equal(out, exp), (nctas, off)
    assert acc.L.libsmm_acc_b200_memset_zero_trickle(d.ptr, 8, 64, 4, acc.s) == -2  # misaligned offset
    d.free()
