"""Property test of the run-aligned chunk partition of the FP64 stack kernel (dbcsr_b200/csrc/smm_dmma.cuh: warp_chunk() +
FLAG_ALIGN_RUNS), restated line by line in Python: for any stack, any split (one-wave / fixed chunk / balanced) the warps' chunks
[boundary(n0), boundary(n1)) tile [0, S) exactly, and a run of equal c_first is only split when it reaches ALIGN_LOOKAHEAD - 1 = 3
or more entries beyond the nominal boundary (round 1 looked 30 entries ahead: bad for stacks with long runs, see smm_dmma.cuh)."""
import numpy as np
import pytest

LOOKAHEAD = 4  # ALIGN_LOOKAHEAD of smm_dmma.cuh


def warp_chunk(gw, chunk, extra, S):
    if extra < 0:
        e0 = min(gw * chunk, S)
        e1 = min(e0 + chunk, S)
    else:
        e0 = min(gw * chunk + min(gw, extra), S)
        e1 = min(e0 + chunk + (1 if gw < extra else 0), S)
    return e0, e1


def aligned(c, n0, n1, S):
    """the kernel's two ballots: lanes 0..LOOKAHEAD-1 look at entries n0.. (start) and lanes 1..LOOKAHEAD at entries n1.. (end)"""
    e0, e1 = n0, n1
    if n0 > 0:
        c_prev = c[n0 - 1]
        for lane in range(LOOKAHEAD):
            if n0 + lane >= S or c[n0 + lane] != c_prev:
                e0 = n0 + lane
                break
    if n1 < S:
        w0 = c[n1 - 1]
        for lane in range(1, LOOKAHEAD + 1):
            if n1 - 1 + lane >= S or c[n1 - 1 + lane] != w0:
                e1 = n1 - 1 + lane
                break
    return e0, e1


@pytest.mark.parametrize("seed", range(6))
def test_aligned_chunks_tile_the_stack(seed):
    rng = np.random.default_rng(seed)
    for _ in range(60):
        S = int(rng.integers(1, 3000))
        mean_run = float(rng.choice([1.0, 1.7, 4.0, 12.0, 45.0, 400.0]))
        runs = np.maximum(1, rng.geometric(1.0 / mean_run, size=S))
        c = np.repeat(np.arange(runs.size), runs)[:S] * 529 + 1
        if rng.random() < 0.2:
            c = rng.permutation(c)  # unsorted stacks are legal too
        wpc = int(rng.choice([2, 4, 8]))
        mode = rng.integers(0, 3)
        if mode == 0:    # one wave
            grid = min(int(rng.integers(1, 900)), (S + wpc * 4 - 1) // (wpc * 4))
        else:            # fixed entries per warp
            grid = (S + wpc * int(rng.integers(1, 41)) - 1) // (wpc * int(rng.integers(1, 41)))
        grid = max(grid, 1)
        warps = grid * wpc
        if mode == 2:
            chunk, extra = S // warps, S % warps
        else:
            chunk, extra = (S + warps - 1) // warps, -1
        covered = np.zeros(S, dtype=np.int32)
        prev_end = 0
        for gw in range(warps):
            n0, n1 = warp_chunk(gw, chunk, extra, S)
            if n0 >= n1:
                continue
            e0, e1 = aligned(c, n0, n1, S)
            if e0 >= e1:
                continue
            assert e0 == prev_end, (S, gw, e0, prev_end)  # consecutive warps meet exactly
            covered[e0:e1] += 1
            prev_end = e1
            # a run is only split at the chunk end when it reaches LOOKAHEAD - 1 or more entries beyond the nominal boundary
            if e1 < S and c[e1] == c[e1 - 1]:
                k = e1
                while k < S and c[k] == c[e1]:
                    k += 1
                assert e1 == n1 and k - n1 >= LOOKAHEAD, (S, gw, n1, e1, k)
        assert prev_end == S and np.all(covered == 1)
