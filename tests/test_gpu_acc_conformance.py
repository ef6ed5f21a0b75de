"""Runs the reference's UNMODIFIED acc-interface conformance test (tests/dbcsr_acc_test.c, built by oracle/Makefile against
the reference's own acc.h and linked with libdbcsr_acc_b200.so) on the GPU: streams, events (OpenMP-parallel create/destroy,
unrecorded events query as occurred), pinned/device memory, memset_zero + d2h checked byte-wise."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "dbcsr_acc_test")


def test_reference_acc_conformance_binary():
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/dbcsr_acc_test not built (needs /root/reference at build time)")
    for nthreads in ("1", "4", "8"):
        env = dict(os.environ, OMP_NUM_THREADS=nthreads)
        r = subprocess.run([BIN, "0", nthreads], env=env, capture_output=True, timeout=120)
        assert r.returncode == 0, (nthreads, r.stdout[-2000:], r.stderr[-2000:])
