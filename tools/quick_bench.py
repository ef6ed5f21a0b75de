"""Quick kernel-only probe (not the bench contract): the reference's timing problem (libsmm_acc_benchmark.cpp:36-44:
16005-entry stack, 10000 A/B blocks, 1000 C blocks) plus an HBM-sized variant, per shape.  Wall-clock around stream sync."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from dbcsr_b200 import lib as acclib
from oracle import oracle as orc


def run(acc, m, n, k, n_a, n_b, n_c, S, reps=20):
    rng = np.random.default_rng(0)
    a, b = rng.random(n_a * m * k), rng.random(n_b * k * n)
    stack = np.empty(3 * S, dtype=np.int32)
    orc.srand(1)
    orc.lib().orc_stack_init(stack, S, n_c, n_a, n_b, m, n, k)
    d_a, d_b, d_s = acc.to_device(a, acc.s), acc.to_device(b, acc.s), acc.to_device(stack, acc.s)
    d_c = acc.dev_alloc(n_c * m * n * 8)
    acc.memset_zero(d_c, acc.s)
    for _ in range(3):
        acc.process(None, d_s.ptr, S, d_a.ptr, d_b.ptr, d_c.ptr, m, n, k, True, acc.s, acc.s)
    acc.stream_sync(acc.s)
    t0 = time.perf_counter()
    for _ in range(reps):
        acc.process(None, d_s.ptr, S, d_a.ptr, d_b.ptr, d_c.ptr, m, n, k, True, acc.s, acc.s)
    acc.stream_sync(acc.s)
    dt = (time.perf_counter() - t0) / reps
    for d in (d_a, d_b, d_s, d_c):
        d.free()
    return 2.0 * m * n * k * S / dt * 1e-9, dt * 1e6


if __name__ == "__main__":
    acc = acclib.Acc(0)
    acc.s = acc.stream_create("qb", 0)
    for (m, n, k) in [(23, 23, 23), (5, 5, 5), (13, 13, 13), (26, 26, 26), (32, 32, 32)]:
        g1, t1 = run(acc, m, n, k, 10000, 10000, 1000, 16005)
        g2, t2 = run(acc, m, n, k, 100000, 100000, 3000, 30000)
        print("%2dx%2dx%2d  ref-timing-problem(L2): %8.1f GFLOP/s (%.1f us)   30000-stack/1e5 blocks: %8.1f GFLOP/s (%.1f us)"
              % (m, n, k, g1, t1, g2, t2), flush=True)
