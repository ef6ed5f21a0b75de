"""Pins the oracle (oracle/dbcsr_oracle.c) against the reference's own known answers:

* the nine golden checksum pairs of the reference's tests/inputs/*.perf (tests/golden/perf_golden.json), which
  exercise the dlarnv-based random-matrix recipe, the block product and dbcsr_checksum end to end;
* LAPACK dlarnv (scipy's bundled OpenBLAS = the third-party BLAS/LAPACK the reference links) for orc_dlarnv1;
* the reference's own C++ checker functions compiled from /root/reference into oracle/_ref (when shipped).
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "perf_golden.json")))


def test_dlarnv_matches_lapack():
    if orc.openblas() is None:
        pytest.skip("no scipy OpenBLAS")
    rng = np.random.default_rng(0)
    for n in (1, 5, 64, 65, 128, 129, 529, 1000):
        seed = [int(v) for v in rng.integers(0, 4096, 4)]
        seed[3] |= 1
        x_ref, seed_ref = orc.lapack_dlarnv(seed, n)
        s = np.array(seed, dtype=np.int32)
        x = np.empty(n)
        orc.lib().orc_dlarnv1(s, n, x)
        assert np.array_equal(x, x_ref)
        assert np.array_equal(s, seed_ref)


def test_set_larnv_seed_constraints():
    s = np.zeros(4, dtype=np.int32)
    for (r, nr, c, nc, v) in [(1, 200, 1, 200, 12341314), (200, 200, 200, 200, 12341316), (7, 42, 3, 42, 12341315)]:
        orc.lib().orc_set_larnv_seed(r, nr, c, nc, v, s)
        assert all(0 <= int(t) < 4096 for t in s) and s[3] % 2 == 1
    # hand-evaluated from src/utils/dbcsr_blas_operations.F:45-50 for (irow=7,nrow=42,icol=3,ncol=42,ival=12341314)
    orc.lib().orc_set_larnv_seed(7, 42, 3, 42, 12341314, s)
    mp = ((7 - 1 + 3 * 42) * (1 + 12341314 % 65536)) * 2 + 1
    exp = [0, 0, 0, mp % 4096]
    mp //= 4096
    exp[2] = (mp ^ 3541) % 4096
    mp //= 4096
    exp[1] = (mp ^ 1153) % 4096
    mp //= 4096
    exp[0] = (mp ^ 2029) % 4096
    assert list(s) == exp


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_perf_golden_checksums(case):
    """generator -> C_out = op(A) B + C_in -> dbcsr_checksum reproduces the stored references (rel. 1e-11)."""
    assert case["transb"] == "N" and case["sym"] == ["N", "N", "N"] and case["alpha"] == [1.0, 0.0] and case["beta"] == [1.0, 0.0]
    sizes_m = orc.random_block_sizes(case["M"], case["bs_m"])
    sizes_n = orc.random_block_sizes(case["N"], case["bs_n"])
    sizes_k = orc.random_block_sizes(case["K"], case["bs_k"])
    ta = case["transa"] == "T"
    # generation order C, A, B with randmat_counter 12341313+1,+2,+3 (tests/dbcsr_performance_multiply.F:271,373-411)
    C = orc.random_matrix(sizes_m, sizes_n, case["sparsity"][2], 12341314)
    A = orc.random_matrix(sizes_k, sizes_m, case["sparsity"][0], 12341315) if ta else \
        orc.random_matrix(sizes_m, sizes_k, case["sparsity"][0], 12341315)
    B = orc.random_matrix(sizes_k, sizes_n, case["sparsity"][1], 12341316)
    Cout = orc.multiply_blocks(A, B, C, transa=ta)
    cs, cs_pos = Cout.checksum(), Cout.checksum(pos=True)
    thr = max(case["threshold"], 1e-11)
    assert abs(cs / case["checksum"] - 1.0) <= thr, (cs, case["checksum"])
    assert abs(cs_pos / case["checksum_pos"] - 1.0) <= thr, (cs_pos, case["checksum_pos"])


def test_restatement_vs_reference_checker():
    """orc_* == the reference's own matInit/stackInit/stackCalc/stackTransp/checkSum compiled from its sources."""
    R = orc.ref()
    if R is None:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    L = orc.lib()
    for (m, n, k) in [(23, 23, 23), (5, 13, 7), (4, 4, 4), (32, 26, 13), (1, 1, 1)]:
        n_a, n_b, n_c, n_stack = 100, 100, 10, 100
        a1, a2 = np.empty(n_a * m * k), np.empty(n_a * m * k)
        b1, b2 = np.empty(n_b * k * n), np.empty(n_b * k * n)
        L.orc_mat_init(a1, n_a, m, k, 42)
        R.ref_matInit(a2, n_a, m, k, 42)
        L.orc_mat_init(b1, n_b, k, n, 24)
        R.ref_matInit(b2, n_b, k, n, 24)
        assert np.array_equal(a1, a2) and np.array_equal(b1, b2)
        s1, s2 = np.empty(3 * n_stack, dtype=np.int32), np.empty(3 * n_stack, dtype=np.int32)
        orc.srand(1)
        L.orc_stack_init(s1, n_stack, n_c, n_a, n_b, m, n, k)
        orc.srand(1)
        R.ref_stackInit(s2, n_stack, n_c, n_a, n_b, m, n, k)
        assert np.array_equal(s1, s2)
        c1, c2 = np.zeros(n_c * m * n), np.zeros(n_c * m * n)
        L.orc_stack_calc(s1, n_stack, c1, a1, b1, m, n, k)
        R.ref_stackCalc(s2, n_stack, c2, a2, b2, m, n, k)
        assert np.array_equal(c1, c2)
        assert L.orc_checksum(c1, n_c, m, n) == R.ref_checkSum(c2, n_c, m, n)
        # transpose: reference is out of place (mat -> mat_trs), ours in place
        st = (np.arange(n_a, dtype=np.int32) * (m * k)).astype(np.int32)
        t_ref = np.zeros_like(a2)
        R.ref_stackTransp(st, n_a, a2, t_ref, m, k)
        t = a1.copy()
        L.orc_transpose(st, n_a, t, m, k)
        assert np.array_equal(t, t_ref)
        assert L.orc_checksum_transp(t, n_a, m, k) == R.ref_checkSumTransp(t_ref, n_a, m, k)


def test_host_stack_equals_device_stack_semantics():
    """blas_process_mm_stack (B normal) on host stacks == stackCalc on the same stack with B transposed in place."""
    rng = np.random.default_rng(1)
    m, n, k = 7, 5, 9
    na = nb = 20
    nc = 6
    a = rng.random(na * m * k)
    b = rng.random(nb * k * n)
    S = 50
    host = np.zeros((S, 7), dtype=np.int32)
    host[:, 0], host[:, 1], host[:, 2] = m, n, k
    host[:, 3] = rng.integers(0, na, S) * m * k + 1
    host[:, 4] = rng.integers(0, nb, S) * k * n + 1
    cb = np.sort(rng.integers(0, nc, S))
    host[:, 5] = cb * m * n + 1
    host[:, 6] = cb + 1
    c_host = orc.host_stack(host, a, b, np.zeros(nc * m * n))
    bt = b.copy()
    orc.transpose_blocks(np.arange(nb, dtype=np.int32) * k * n, bt, k, n)
    c_dev = orc.stack_calc(host[:, 3:6].copy(), np.zeros(nc * m * n), a, bt, m, n, k)
    assert np.allclose(c_host, c_dev, rtol=1e-13, atol=0)
    if orc.dgemm_ptr() is not None:
        c_blas = orc.host_stack(host, a, b, np.zeros(nc * m * n), use_blas=True)
        assert np.linalg.norm(c_blas - c_host) / np.linalg.norm(c_host) < 1e-14


def test_norms():
    rng = np.random.default_rng(2)
    mat = rng.random(1000)
    offs = np.array([0, 10, 500], dtype=np.int32)
    ne = np.array([10, 25, 500], dtype=np.int32)
    out = orc.norms(mat, offs, ne)
    exp = np.array([np.sum(mat[o:o + e] ** 2) for o, e in zip(offs, ne)], dtype=np.float32)
    assert np.allclose(out, exp, rtol=1e-6)


def test_h2o_statistics_from_reference_docs():
    """Known answer from the reference's documentation (docs/guide/3-developer-guide/4-performance/1-insights.md:21-35): the DBCSR
    STATISTICS table of `dbcsr_perf tests/inputs/test_H2O.perf` (2208^3, 23x23 blocks, sparsity 0.2, 50 multiplications, 1 rank):
    flops 23x23x23 = 687272462200, matmuls total = 28243300, 1600 stacks of average size 17652.1 (= 32 threads x 1 stack x 50).
    Reproduces the block patterns of the dlarnv/geometric-skipping generator, the product enumeration of the stack builder, the
    flop count (src/mm/dbcsr_mm_csr.F:350) and the one-stack-per-thread purge behaviour."""
    from dbcsr_b200 import host

    sizes = orc.random_block_sizes(2208, [1, 23])
    assert len(sizes) == 96
    A = orc.random_matrix(sizes, sizes, 0.2, 12341315)  # generation order C, A, B after the seed reset
    B = orc.random_matrix(sizes, sizes, 0.2, 12341316)
    eng = host.Engine(sizes, sizes, sizes, nthreads=32, mode=host.RECORD)
    eng.multiply(np.array(A.index_list(), dtype=np.int32), None, np.array(B.index_list(), dtype=np.int32), None)
    stacks = eng.stacks()
    matmuls = sum(s["host"].shape[0] for s in stacks)
    assert matmuls * 50 == 28243300
    assert eng.flop() * 50 == 687272462200
    assert len(stacks) * 50 == 1600
    assert abs(matmuls / len(stacks) - 17652.1) < 0.05
    assert all(s["m"] == 23 and s["n"] == 23 and s["k"] == 23 and s["defined_mnk"] for s in stacks)
    # the same numbers through the engine's own STATISTICS table (dbcsr_mm_sched's counters, merged over the 32 threads), after
    # the 50 multiplications of the perf run: one (m,n,k) row "flops 23 x 23 x 23" and the totals of the documentation
    for _ in range(49):
        eng.reset()
        eng.multiply(np.array(A.index_list(), dtype=np.int32), None, np.array(B.index_list(), dtype=np.int32), None)
    rows, totals = eng.stats()
    assert len(rows) == 1 and (rows[0]["m"], rows[0]["n"], rows[0]["k"]) == (23, 23, 23)
    assert rows[0]["flop"] == 687272462200 and rows[0]["entries"] == 28243300 and rows[0]["stacks"] == 1600
    assert totals == dict(flop=687272462200, entries=28243300, stacks=1600) and rows[0]["stacks_untuned"] == 0
    eng.close()
