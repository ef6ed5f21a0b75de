"""Digest of the kernel-timeline records written by the TRACE variants of the FP64 stack kernels (tools/kbench.c spec `...:t`,
record layout in dbcsr_b200/csrc/smm_dmma.cuh): per-entry copy latency, compute-phase length, inter-entry gap, flush cost,
launch overlap and the SM clock (cycles / globaltimer).  Usage: python tools/trace_analyze.py gpurun_out/trace_*.bin"""
import sys

import numpy as np


def analyze(path, slot=1):
    t = np.fromfile(path, dtype=np.uint64).reshape(3, 4096, 128).astype(np.int64)
    r = t[slot]
    r = r[r[:, 2] != 0]
    nent = r[:, 0] >> 32
    n = min(int(np.median(nent)), 30)
    rr = r[nent >= n]
    E = np.arange(n)
    tI, tA, tD = rr[:, 4 + 4 * E], rr[:, 5 + 4 * E], rr[:, 6 + 4 * E]
    L, C, gap = tA - tI, tD - tA, tI[:, 1:] - tD[:, :-1]
    life_c, life_g = rr[:, 126] - rr[:, 2], rr[:, 127] - rr[:, 1]
    g0, g1 = r[:, 1], r[:, 127]
    out = {
        "file": path, "warps": int(len(r)), "entries_per_warp_median": n, "sm_clock_ghz": float(life_c.sum() / life_g.sum()),
        "start_to_first_issue_cycles": float((tI[:, 0] - rr[:, 2]).mean()),
        "copy_latency_cycles_mean_per_entry": np.round(L.mean(0)).astype(int).tolist(),
        "compute_phase_cycles_mean_per_entry": np.round(C.mean(0)).astype(int).tolist(),
        "compute_phase_percentiles_1_5_50_95": [int(np.percentile(C, p)) for p in (1, 5, 50, 95)],
        "gap_cycles_mean_per_entry": np.round(gap.mean(0)).astype(int).tolist(),
        "flushes_per_warp": float(rr[:, 124].mean()), "cycles_per_flush": float((rr[:, 125] / np.maximum(rr[:, 124], 1)).mean()),
        "warp_lifetime_cycles": float(life_c.mean()), "cycles_per_entry_per_warp": float((life_c / nent[nent >= n]).mean()),
        "launch_span_us": float((g1.max() - g0.min()) / 1e3), "start_spread_us": float((g0.max() - g0.min()) / 1e3),
        "launch_period_us": [float((t[s + 1][t[s + 1][:, 2] != 0][:, 1].min() - t[s][t[s][:, 2] != 0][:, 1].min()) / 1e3) for s in range(2)],
    }
    return out


if __name__ == "__main__":
    import json

    for p in sys.argv[1:]:
        print(json.dumps(analyze(p)))
