"""GPU twin of tests/test_dbcsr_multiply.py: the reference's dbcsr_multiply unit-test cases (tests/dbcsr_unittest1.F:95-330,
driver tests/dbcsr_test_multiply.F) through the dbcsr_multiply mirror with the DEVICE backend: panels on the GPU, stacks by the
host engine, drained by libsmm_acc_process (tuned DMMA kernels for 5/13/23 blocks, run-time-shape / generic kernels and
inhomogeneous stacks for the 1..4-sized blocks of the reference's cases).  Criterion of dbcsr_check_multiply:
||C_dbcsr - C_dense||_oo / ((||A||_oo + ||B||_oo + ||C_in||_oo) * n * eps) <= 10."""
import os
import zlib

import numpy as np
import pytest

from dbcsr_b200 import dbcsr as D

from dbcsr_multiply_cases import UNITTEST1_CASES, golden_cases, random_matrix, run_case, run_golden_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def backend():
    from dbcsr_b200 import lib as acclib

    acc = acclib.Acc(0)
    yield D.DeviceBackend(acc, nthreads=2)
    acc.finalize()


@pytest.fixture(scope="module")
def backend_dev(backend):
    """the same operator path with the stacks built on the device (DBCSR_B200_DEVICE_BUILD)"""
    return D.DeviceBackend(backend.acc, nthreads=2, device_build=True)


@pytest.mark.parametrize("case", UNITTEST1_CASES[::3], ids=[c[0] for c in UNITTEST1_CASES[::3]])
def test_dbcsr_multiply_unittest_cases_device_builder(case, backend_dev):
    """Every third parameter set of the reference's unit tests (alpha/beta, retain_sparsity, limits, symmetry triples, transposes)
    with the device-side stack builder under the operator."""
    rng = np.random.default_rng(zlib.crc32(case[0].encode()))
    n = 0
    for desc, eps_norm, flop in run_case(case, backend_dev, rng):
        assert eps_norm <= 10.0, (desc, eps_norm)
        n += 1
    assert n >= 1, n


@pytest.mark.parametrize("case", golden_cases()[::2], ids=[c["name"] for c in golden_cases()[::2]])
def test_perf_golden_checksums_device_builder(case, backend_dev):
    """The reference's stored checksums (tests/inputs/*.perf) with the stacks built on the device."""
    cs, cs_pos = run_golden_case(case, backend_dev)
    thr = max(case["threshold"], 1e-11)
    assert abs(cs / case["checksum"] - 1.0) <= thr, (cs, case["checksum"])
    assert abs(cs_pos / case["checksum_pos"] - 1.0) <= thr, (cs_pos, case["checksum_pos"])


@pytest.mark.parametrize("case", UNITTEST1_CASES, ids=[c[0] for c in UNITTEST1_CASES])
def test_dbcsr_multiply_unittest_cases_on_device(case, backend):
    rng = np.random.default_rng(zlib.crc32(case[0].encode()))  # deterministic per case
    n = 0
    for desc, eps_norm, flop in run_case(case, backend, rng):
        assert eps_norm <= 10.0, (desc, eps_norm)
        n += 1
    assert n >= 1, n


def test_filter_eps_on_device(backend):
    """filter_eps through the device path: device norms (c_calculate_norms), on-the-fly filter in the host engine, final filter +
    compaction of C on the device before the download."""
    from oracle import oracle as orc

    rng = np.random.default_rng(5)
    sizes = orc.random_block_sizes(92, [1, 5, 1, 13, 1, 23])
    a = random_matrix("A", sizes, sizes, 0.6, "N", rng)
    b = random_matrix("B", sizes, sizes, 0.6, "N", rng)
    for i in range(0, a.nblks, 3):
        a.block(i)[...] *= 1e-7
    exact = a.to_dense() @ b.to_dense()
    c = D.DbcsrMatrix("C", sizes, sizes)
    eps = 1e-4
    D.dbcsr_multiply("N", "N", 1.0, a, b, 0.0, c, filter_eps=eps, backend=backend)
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


def test_device_finalize_gives_bcsr_order_and_same_blocks(backend):
    """dbcsr_b200_engine_finalize_c: after it every thread's index is in BCSR order with compact offsets, and the merged product
    equals the one merged on the host from the first-touch-ordered work matrices (bit for bit)."""
    from dbcsr_b200 import host, workload
    from dbcsr_b200.multiply import DeviceMultiply

    rng = np.random.default_rng(9)
    bs = workload.block_sizes(40, [5, 13, 23], rng)
    A = workload.random_panel(bs, bs, 0.3, rng)
    B = workload.random_panel(bs, bs, 0.3, rng)
    dm = DeviceMultiply(backend.acc, bs, bs, bs, A.data.size, B.data.size, B.nblks, nthreads=3, cfg=host.default_cfg(mm_stack_size=300))
    try:
        dm.upload_panels(A.data, B.data, B.list3())
        dm.multiply(A.list3(), B.list3())
        raw = dm.download_c()
        ref = D.dbcsr_finalize(bs, bs, [(r, c, p, np.array(d, copy=True)) for (r, c, p, d) in raw.parts])
        dm.finalize_c()
        fin = dm.download_c()
        for rows, cols, blk_p, data in fin.parts:
            key = rows.astype(np.int64) * 100000 + cols
            assert np.all(key[1:] > key[:-1])
            nze = bs[rows - 1].astype(np.int64) * bs[cols - 1]
            assert np.array_equal(blk_p, 1 + np.concatenate([[0], np.cumsum(nze)[:-1]])) and data.size == int(nze.sum())
        got = D.dbcsr_finalize(bs, bs, [(r, c, p, np.array(d, copy=True)) for (r, c, p, d) in fin.parts])
        assert np.array_equal(got.row_p, ref.row_p) and np.array_equal(got.col_i, ref.col_i) and np.array_equal(got.blk_p, ref.blk_p)
        assert np.array_equal(got.data, ref.data)
        # with the final filter: survivors only, still sorted
        dm.multiply(A.list3(), B.list3())
        norms2 = {k: float(np.dot(b.T.reshape(-1), b.T.reshape(-1))) for k, b in ref.blocks().items()}
        eps = float(np.sqrt(np.median(list(norms2.values())))) * 1.0000001  # about half of the blocks survive
        dm.finalize_c(filter_eps=eps)
        flt = dm.download_c()
        kept = D.dbcsr_finalize(bs, bs, [(r, c, p, np.array(d, copy=True)) for (r, c, p, d) in flt.parts])
        want = {k for k, v in norms2.items() if v >= eps * eps * (1 + 1e-12)}
        borderline = {k for k, v in norms2.items() if abs(v - eps * eps) <= 1e-9 * eps * eps}
        assert set(kept.blocks()) - borderline == want - borderline
        assert 0 < len(want) < ref.nblks
    finally:
        dm.close()


@pytest.mark.parametrize("case", golden_cases(), ids=[c["name"] for c in golden_cases()])
def test_perf_golden_checksums_on_device(case, backend):
    """The device path against the reference's STORED results: the nine golden checksum pairs of tests/inputs/*.perf reproduced
    with the multiply running on the GPU (dbcsr_multiply -> host engine -> libsmm_acc_process: 5x5x5 DMMA kernel, run-time-shape
    and generic kernels for the other block sizes), rel. 1e-11 like the reference's own check."""
    cs, cs_pos = run_golden_case(case, backend)
    thr = max(case["threshold"], 1e-11)
    assert abs(cs / case["checksum"] - 1.0) <= thr, (cs, case["checksum"])
    assert abs(cs_pos / case["checksum_pos"] - 1.0) <= thr, (cs_pos, case["checksum_pos"])


@pytest.mark.parametrize("bigdmma", [1, 0])
@pytest.mark.parametrize("mnk", [(33, 33, 33), (45, 67, 78), (80, 80, 80), (64, 40, 72), (40, 17, 7), (9, 80, 33)])
def test_cooperative_dmma_kernel_for_blocks_33_to_80(backend, mnk, bigdmma):
    """smm_dmma_big.cuh (the default for blocks with a dimension in 33..80) and the scalar generic kernel it replaced
    ("bigdmma" = 0): same contract (return code 10 = untuned, exact on integer inputs), incl. runs, unsorted stacks and odd
    alignment."""
    from oracle import oracle as orc
    from test_gpu_smm import run_process

    acc = backend.acc
    if not hasattr(acc, "s"):
        acc.s = acc.stream_create("big", 0)
    m, n, k = mnk
    rng = np.random.default_rng(4)
    n_a = n_b = 60
    a = rng.integers(0, 4, n_a * m * k).astype(np.float64)
    b = rng.integers(0, 4, n_b * k * n).astype(np.float64)
    saved = acc.get_tunable("bigdmma")
    acc.set_tunable("bigdmma", bigdmma)
    try:
        for S, n_c, shuffle, pad in [(1, 1, False, 0), (37, 5, False, 1), (500, 40, False, 0), (300, 7, True, 1)]:
            stack = np.zeros((S, 3), dtype=np.int32)
            stack[:, 0] = rng.integers(0, n_a, S) * m * k + 1
            stack[:, 1] = rng.integers(0, n_b, S) * k * n + 1
            stack[:, 2] = np.sort(rng.integers(0, n_c, S)) * m * n + 1
            if shuffle:
                stack = stack[rng.permutation(S)]
            c_ref = orc.stack_calc(stack, np.zeros(n_c * m * n), a, b, m, n, k)
            rc, c = run_process(acc, stack, a, b, n_c * m * n, m, n, k, pad_elems=pad)
            assert rc == 10
            assert np.array_equal(c, c_ref), (mnk, S, float(np.abs(c - c_ref).max()))
    finally:
        acc.set_tunable("bigdmma", saved)
