"""Multi-process Cannon multiply on real GPUs, checked block by block against the oracle.
The reference tests every distributed multiply against dense DGEMM with 2 MPI ranks (tests/CMakeLists.txt:130-137,
tests/dbcsr_test_multiply.F:753-759); here 2, 4 and 8 ranks = 1x2, 2x2 and 2x4 Cannon grids (2x4: four virtual k-slices on a non-square grid), engine path and replay path, 23x23 and
mixed block sizes, plus one multiply whose input is DISTRIBUTED (every rank holds only its blocks under an unrelated 2-d block
distribution; `cannon.make_images` = the reference's make_images moves them to the home panels by an all-to-all).  One rank per GPU over NCCL when the box has enough GPUs; otherwise the ranks share GPU 0 (gloo set-up
collectives, CUDA IPC peer pull) so that the distributed path is exercised on a single-GPU box too."""
import os
import re
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_cannon_blocks_match_oracle(world):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import tempfile

    outdir = tempfile.mkdtemp(prefix="cannon_test_")
    env = dict(os.environ, OMP_NUM_THREADS="2", CANNON_TEST_OUT=outdir)  # every rank writes its report to a file: stdout lines of
    env.pop("DBCSR_B200_EXCHANGE", None)                                # concurrently printing ranks interleave
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "cannon_gpu_worker.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)
    import json

    reports = []
    for r in range(world):
        f = os.path.join(outdir, "rank%d.json" % r)
        if os.path.exists(f):
            reports.append(json.load(open(f)))
    assert out.returncode == 0 and len(reports) == world, "rc %d, %d reports\nstdout:\n%s\nstderr:\n%s" % (out.returncode, len(reports), out.stdout[-3000:],
                                                                                                        out.stderr[-3000:])
    for rep in reports:
        assert len(rep["cases"]) == 5  # cfg2 / cfg3 x (engine, replay) + cfg2 with distributed input through make_images
        for case in rep["cases"]:
            assert case["blocks"] > 0 and case["worst_rel_err"] <= 1e-10
