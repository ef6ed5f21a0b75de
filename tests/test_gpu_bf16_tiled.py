"""Tiled BF16 SpGEMM (libsmm_acc_b200_bf16_spgemm, tcgen05 with TMEM-resident C tiles; BASELINE.json config 4): parity against
the FP64 oracle (orc.multiply_blocks = plain block products of the reference's CPU arithmetic).
  * tolerance of BASELINE.json north_star: relative Frobenius error <= 1e-3 against the oracle on the UNROUNDED inputs;
  * sharper: against the oracle on the bf16-ROUNDED inputs only FP32 accumulation differs, which pins the operand layout, the
    descriptors, the presence-map handling (absent A slots read as zeros, absent B blocks are skipped) and the C tile stores.
Edge cases: ragged tile edges (block counts that are no multiple of 5 / 16), empty block rows / columns / k blocks, one block,
occupations from 2 % to 100 %, block sizes 1 ... 32 incl. rectangular."""
import numpy as np
import pytest

from dbcsr_b200 import workload
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def acc():
    from dbcsr_b200 import lib as acclib

    a = acclib.Acc(0)
    a.s = a.stream_create("bf16 tiled", 0)
    yield a
    a.stream_destroy(a.s)
    a.finalize()


def bf16_round(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) >> 16 << 16
    return u.astype(np.uint32).view(np.float32).astype(np.float64)


def oracle_dense_blocks(A, B, data_a, data_b, m, n):
    """(nrb, ncb, n, m) array of the oracle's C blocks (zeros where the pattern product is empty)."""
    C = orc.multiply_blocks(orc.BlockMatrix(A.row_sizes, A.col_sizes, A.rows, A.cols, data=data_a),
                            orc.BlockMatrix(B.row_sizes, B.col_sizes, B.rows, B.cols, data=data_b))
    out = np.zeros((A.row_sizes.size, B.col_sizes.size, n, m))
    for r, c, o in zip(C.rows, C.cols, C.offsets):
        out[r - 1, c - 1] = C.data[o:o + m * n].reshape(n, m)
    return out


def run_case(acc, nrb, ncb, nkb, m, n, k, occ_a, occ_b, seed, knock_out=False):
    from dbcsr_b200.bf16 import Bf16SpGemm

    rng = np.random.default_rng(seed)
    A = workload.random_panel(np.full(nrb, m, np.int32), np.full(nkb, k, np.int32), occ_a, rng)
    B = workload.random_panel(np.full(nkb, k, np.int32), np.full(ncb, n, np.int32), occ_b, rng)
    if knock_out and A.nblks > 4 and B.nblks > 4:
        # an empty block row of A, an empty block column of B and an empty k block
        keep = (A.rows != 2) & (A.cols != 3)
        A = workload.Panel(A.row_sizes, A.col_sizes, A.rows[keep], A.cols[keep], rng=rng)
        keep = (B.cols != 1) & (B.rows != 3)
        B = workload.Panel(B.row_sizes, B.col_sizes, B.rows[keep], B.cols[keep], rng=rng)
    mm = Bf16SpGemm(acc, A, B, acc.s)
    # poison C: the kernel overwrites every block it is asked for (also the ones without any contribution)
    acc.h2d(np.full(max(mm.c_elems, 1), np.float32(7.5)), mm.d_c, acc.s)
    mm.run()
    got = mm.result().astype(np.float64)
    # second run into the same buffer: identical (no accumulation onto old contents)
    mm.run()
    again = mm.result().astype(np.float64)
    mm.close()
    assert np.array_equal(got, again)
    ref = oracle_dense_blocks(A, B, A.data, B.data, m, n)
    ref_r = oracle_dense_blocks(A, B, bf16_round(A.data), bf16_round(B.data), m, n)
    den = max(np.linalg.norm(ref), 1e-300)
    zero_blocks = np.abs(ref).sum(axis=(2, 3)) == 0  # structure: blocks without any contribution are exact zeros
    assert np.all(got[zero_blocks] == 0.0)
    err = np.linalg.norm(got - ref) / den
    err_r = np.linalg.norm(got - ref_r) / max(np.linalg.norm(ref_r), 1e-300)
    assert mm.products == int(sum(np.sum(A.cols == kk) * np.sum(B.rows == kk) for kk in range(1, nkb + 1)))
    return err, err_r


# kernel modes (run-time tunables, dbcsr_b200/csrc/smm_tune.h): one MMA per existing B block / per run of adjacent blocks, A operand
# read from shared memory by every MMA / staged once per k block in TMEM (tcgen05.cp)
# bf16_plan=1 (default): the planned kernel (smm_bf16_plan.cuh: copy commands and MMA runs derived once per multiply by bt_plan_kernel,
# accumulators zeroed by the epilogue); bf16_plan=0: the first tiled kernel, which derives everything from the presence maps per k block
MODES = [dict(bf16_plan=1, bf16_merge=1, bf16_a_tmem=0), dict(bf16_plan=1, bf16_merge=1, bf16_a_tmem=1),
         dict(bf16_plan=0, bf16_merge=1, bf16_a_tmem=0), dict(bf16_plan=0, bf16_merge=0, bf16_a_tmem=0),
         dict(bf16_plan=0, bf16_merge=1, bf16_a_tmem=1), dict(bf16_plan=0, bf16_merge=0, bf16_a_tmem=1)]


@pytest.fixture(params=MODES, ids=lambda m: "plan%d_merge%d_atmem%d" % (m["bf16_plan"], m["bf16_merge"], m["bf16_a_tmem"]))
def mode(acc, request):
    saved = {k: acc.get_tunable(k) for k in request.param}
    for k, v in request.param.items():
        acc.set_tunable(k, v)
    yield request.param
    for k, v in saved.items():
        acc.set_tunable(k, v)


@pytest.mark.parametrize("shape", [(5, 16, 4), (1, 1, 1), (7, 19, 9), (23, 40, 31), (64, 64, 64), (3, 70, 2)])
def test_tiled_bf16_23_blocks(acc, mode, shape):
    nrb, ncb, nkb = shape
    err, err_r = run_case(acc, nrb, ncb, nkb, 23, 23, 23, 0.5, 0.5, seed=11 + nrb)
    assert err <= 1e-3, err      # north_star tolerance for BF16
    assert err_r <= 5e-6, err_r  # only FP32 accumulation differs


@pytest.mark.parametrize("occ", [0.02, 0.1, 1.0])
def test_tiled_bf16_occupations_and_empty_rows(acc, mode, occ):
    err, err_r = run_case(acc, 33, 37, 29, 23, 23, 23, occ, occ, seed=3, knock_out=True)
    assert err <= 1e-3 and err_r <= 5e-6, (err, err_r)


@pytest.mark.parametrize("mnk", [(5, 5, 5), (13, 13, 13), (26, 26, 26), (32, 32, 32), (23, 5, 32), (13, 32, 5), (8, 16, 16), (1, 1, 1), (24, 24, 24)])
def test_tiled_bf16_block_sizes(acc, mode, mnk):
    m, n, k = mnk
    err, err_r = run_case(acc, 21, 35, 17, m, n, k, 0.4, 0.6, seed=5)
    # 1e-3 is the tolerance of the 23x23x23 configuration, where a C element averages the BF16 rounding of 23 x ~250 terms; a C
    # element of 1x1 blocks sums ~4 products, each off by up to 2^-8 relative (two operands rounded to 8 mantissa bits)
    tol = 1e-3 if k >= 13 else 4e-3
    assert err <= tol and err_r <= 5e-6, (mnk, err, err_r)


def test_tiled_bf16_rejects_large_blocks(acc):
    """m, n or k above 32: -10 and nothing is enqueued (C untouched)."""
    from dbcsr_b200 import lib as acclib

    d_c = acc.dev_alloc(4 * 64)
    acc.h2d(np.full(64, np.float32(3.0)), d_c, acc.s)
    rc = acc.L.libsmm_acc_b200_bf16_spgemm(d_c.ptr, d_c.ptr, d_c.ptr, d_c.ptr, d_c.ptr, d_c.ptr, 1, 1, 1, 33, 8, 8, acc.s)
    assert rc == -10
    c = acc.to_host(d_c, (64,), np.float32, acc.s)
    assert np.all(c == 3.0)
    d_c.free()
