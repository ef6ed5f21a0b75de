"""BF16 extension (dbcsr_type_bf16_ext = 9, tcgen05 tensor-core kernel): parity against the FP64 oracle.
  * tolerance of BASELINE.json north_star: relative Frobenius error <= 1e-3 against the FP64 oracle on the UNROUNDED inputs;
  * sharper check: against the FP64 oracle on the bf16-ROUNDED inputs the only differences are FP32 accumulation order (<= 2e-6),
    which pins tile layout, descriptors and run handling exactly."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def acc():
    from dbcsr_b200 import lib as acclib

    a = acclib.Acc(0)
    a.s = a.stream_create("bf16", 0)
    yield a
    a.stream_destroy(a.s)
    a.finalize()


def bf16_round(x):
    """float64 -> nearest-even bfloat16 -> float64 (numpy emulation of the pack kernel's rounding)."""
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) >> 16 << 16
    return u.astype(np.uint32).view(np.float32).astype(np.float64)


def run_bf16(acc, stack3, a, bt, n_c, m, n, k):
    from dbcsr_b200 import lib as acclib

    n_a, n_b = a.size // (m * k), bt.size // (n * k)
    d_a, d_b = acc.to_device(a, acc.s), acc.to_device(bt, acc.s)
    ta, tb = acc.bf16_tile_bytes(m, k), acc.bf16_tile_bytes(n, k)
    p_a, p_b = acc.dev_alloc(n_a * ta), acc.dev_alloc(n_b * tb)
    acc.pack_bf16(d_a.ptr, n_a, m, k, 1, m, p_a.ptr, acc.s)   # A: m x k col-major
    acc.pack_bf16(d_b.ptr, n_b, n, k, 1, n, p_b.ptr, acc.s)   # Bt: n x k col-major
    d_s = acc.to_device(np.ascontiguousarray(stack3, dtype=np.int32), acc.s)
    d_c = acc.dev_alloc(n_c * m * n * 4)
    acc.memset_zero(d_c, acc.s)
    rc = acc.process(None, d_s.ptr, stack3.shape[0], p_a.ptr, p_b.ptr, d_c.ptr, m, n, k, True, acc.s, acc.s,
                     datatype=acclib.DBCSR_TYPE_BF16_EXT)
    c = acc.to_host(d_c, (n_c * m * n,), np.float32, acc.s)
    tiles_a = acc.to_host(p_a, (n_a * ta // 2,), np.uint16, acc.s)
    for d in (d_a, d_b, p_a, p_b, d_s, d_c):
        d.free()
    return rc, c.astype(np.float64), tiles_a


def test_pack_layout_and_rounding(acc):
    m, k = 23, 23
    rng = np.random.default_rng(0)
    n_a = 7
    a = rng.random(n_a * m * k)
    stack = np.array([[1, 1, 1]], dtype=np.int32)
    _, _, tiles = run_bf16(acc, stack, a, rng.random(m * k), 1, m, m, k)
    rg, kg = 3, 3
    tiles = tiles.reshape(n_a, kg, rg, 8, 8)  # [blk][k group][row group][row%8][k%8]
    exp = np.zeros((n_a, kg * 8, rg * 8))     # [blk][kk][row]
    blocks = bf16_round(a).reshape(n_a, k, m)  # col-major: [blk][kk][row]
    exp[:, :k, :m] = blocks
    got = (tiles.astype(np.uint32) << 16).view(np.float32).astype(np.float64)  # [blk][kg][rg][r8][k8]
    got = got.transpose(0, 1, 4, 2, 3).reshape(n_a, kg * 8, rg * 8)            # -> [blk][kk][row]
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("mnk", [(23, 23, 23), (5, 5, 5), (13, 13, 13), (26, 26, 26), (32, 32, 32), (23, 5, 32), (13, 32, 5), (8, 16, 16), (1, 1, 1)])
def test_bf16_process_vs_fp64_oracle(acc, mnk):
    m, n, k = mnk
    rng = np.random.default_rng(5)
    n_a, n_b, n_c, S = 300, 300, 40, 3000
    a, bt = rng.random(n_a * m * k), rng.random(n_b * k * n)
    stack = np.empty(3 * S, dtype=np.int32)
    orc.srand(9)
    orc.lib().orc_stack_init(stack, S, n_c, n_a, n_b, m, n, k)   # C-sorted, runs of ~75 entries
    stack = stack.reshape(-1, 3)
    rc, c, _ = run_bf16(acc, stack, a, bt, n_c, m, n, k)
    assert rc == 0
    c_ref = orc.stack_calc(stack, np.zeros(n_c * m * n), a, bt, m, n, k)
    err = np.linalg.norm(c - c_ref) / np.linalg.norm(c_ref)
    assert err <= 1e-3, err                                    # north_star tolerance for BF16
    c_ref_r = orc.stack_calc(stack, np.zeros(n_c * m * n), bf16_round(a), bf16_round(bt), m, n, k)
    err_r = np.linalg.norm(c - c_ref_r) / np.linalg.norm(c_ref_r)
    assert err_r <= 2e-6, err_r                                # only FP32 accumulation differs


def test_bf16_edge_stacks(acc):
    m = n = k = 23
    rng = np.random.default_rng(6)
    n_a = n_b = 64
    a, bt = rng.random(n_a * m * k), rng.random(n_b * k * n)
    for S, n_c, shuffle in [(1, 1, False), (17, 17, False), (500, 1, False), (2000, 100, True), (30000, 120, False)]:
        stack = np.zeros((S, 3), dtype=np.int32)
        stack[:, 0] = rng.integers(0, n_a, S) * m * k + 1
        stack[:, 1] = rng.integers(0, n_b, S) * k * n + 1
        stack[:, 2] = np.sort(rng.integers(0, n_c, S)) * m * n + 1
        if shuffle:
            stack = stack[rng.permutation(S)]
        rc, c, _ = run_bf16(acc, stack, a, bt, n_c, m, n, k)
        assert rc == 0
        c_ref = orc.stack_calc(stack, np.zeros(n_c * m * n), bf16_round(a), bf16_round(bt), m, n, k)
        assert np.linalg.norm(c - c_ref) / np.linalg.norm(c_ref) <= 5e-6, (S, n_c)
