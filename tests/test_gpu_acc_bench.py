"""Runs the pure-C miniapp tools/acc_bench.c (restatement of the reference's src/acc/acc_bench.c against the drop-in ABI):
pinned/device buffers, transpose + process through C only, validated against a host loop with the CPU path's stack semantics."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tools", "acc_bench")


@pytest.mark.parametrize("args", [["3", "30000", "23", "23", "23"], ["3", "5000", "13", "26", "5"], ["2", "2000", "7", "9", "11"], ["2", "500", "45", "67", "78"]])
def test_acc_bench_miniapp(args):
    if not os.path.exists(BIN):
        pytest.skip("tools/acc_bench not built (run __graft_entry__.build())")
    r = subprocess.run([BIN] + args, capture_output=True, timeout=120, text=True)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    assert "GFLOPS/s" in r.stdout and "max.error" in r.stdout
