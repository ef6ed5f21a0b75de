#!/usr/bin/env python
"""
bench.py -- block-sparse GEMM GFLOP/s of the DBCSR stack-drain hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W                 our arm (C ABI -> sm_100a kernels)
  python bench.py --impl reference --gpus N --steps K --warmup W  the reference's CPU path (oracle port of
                                                                 blas_process_mm_stack_d, src/mm/dbcsr_mm_hostdrv.F:248-282)

Workload at N=1 = BASELINE.json configs[1]: 1000x1000 block grid, 23x23 FP64 blocks, 10 % occupation (1e5 blocks per operand,
423 MB each; 1e7 block products, 2.43e11 flop; C ~1e6 blocks, 4.23 GB), seeded synthetic data.
A step = one whole local multiply drained on the GPU:
  value : stacks pre-built and resident in HBM, A/B/C resident; per step C is zeroed and all ~334 stacks of 30000 entries are
          drained through libsmm_acc_process (stack-kernel only, mirrors src/acc/acc_bench.c:338-345).
  e2e   : same multiply from pinned HOST buffers through the host engine: H2D of both panels, device transpose of the right
          panel, stack building/sorting on the host threads, H2D of every stack, kernels, D2H of C.
At N>1 every rank is one Cannon grid rank (dbcsr_b200/cannon.py): C is sharded over a pr x pc grid, panels move by NCCL
send/recv; strong scaling (total work fixed), value = total flop / max-over-ranks time.
GFLOP/s counts 2*m*n*k per block product (src/mm/dbcsr_mm_csr.F:350); padded tensor-core flops never count.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


METRIC_NAMES = {"cfg2": "block-sparse GEMM GFLOP/s (FP64, 23^3 blocks, 10% occ)",
                "cfg3": "block-sparse GEMM GFLOP/s (FP64, mixed blocks {5,13,23,26,32}, 5% occ)",
                "cfg4": "block-sparse GEMM GFLOP/s (BF16 operands / FP32 accumulate, 23^3 blocks, 50% occ)"}


# ------------------------------------------------------------------------------------------------ helpers
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smmax, power, reasons, trace = [], [], [], set(), []
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smmax.append(float(f[2]))
                power.append(float(f[3]))
                trace.append([float(f[1]), float(f[3])])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        load = [s for s in sm if s > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_min_mhz_under_load": float(min(load)) if load else None,
                "sm_max_mhz": max(smmax) if smmax else None, "power_w_max": max(power) if power else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "trace_every_20ms_sm_mhz_power_w": trace[::max(1, len(trace) // 60)]}


def algorithmic_bytes(stacks):
    """per entry 8(mk+kn) + 12 (stack entry), plus one read+write of the C block per run of equal c_first."""
    total = 0
    for s in stacks:
        dev = s["dev"]
        S = dev.shape[0]
        m, n, k = s["m"], s["n"], s["k"]
        runs = 1 + int(np.count_nonzero(dev[1:, 2] != dev[:-1, 2])) if S else 0
        total += S * (8 * (m * k + k * n) + 12) + runs * 2 * 8 * m * n
    return total


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_reference_sample(w, target_entries, cores=None, n_stacks=3):
    """Times the oracle port of the reference CPU path (one DGEMM('N','N') per stack entry, OpenMP threads own disjoint C rows,
    MM_STACK_SIZE=1000 as in CPU builds) on a bounded sample of the workload's stacks.  Returns (gflops, info)."""
    from dbcsr_b200 import hostbuilder  # host-only library (no accelerator code): DBCSR-order stacks for the CPU arm
    from oracle import oracle as orc

    # all the host cores this process may use -- NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 and the oracle sets
    # its team size explicitly (omp parallel num_threads(cores))
    try:
        ncores = len(os.sched_getaffinity(0))
    except AttributeError:
        ncores = os.cpu_count() or 1
    cores = max(1, min(cores or ncores, 64))
    A, B, bs = w["A"], w["B"], w["m_sizes"]
    t0 = time.perf_counter()
    stacks, datasizes, _ = hostbuilder.record_stacks(bs, bs, bs, A.list3(), B.list3(), nthreads=cores, mm_stack_size=1000, n_stacks=n_stacks)
    t_build = time.perf_counter() - t0
    total_entries = sum(s["host"].shape[0] for s in stacks)
    # per-thread C areas laid out one after the other; sample = the first stacks of every thread up to the entry budget
    bases, base = [], 0
    for t in range(cores):
        bases.append(base)
        base += datasizes[t]
    per_thread_budget = max(1, target_entries // cores)
    taken = [0] * cores
    sel = []
    for s in stacks:
        t = s["thread"]
        if taken[t] < per_thread_budget:
            sel.append(s)
            taken[t] += s["host"].shape[0]
    params = np.concatenate([s["host"] for s in sel]).astype(np.int32)
    ptr = np.zeros(len(sel) + 1, dtype=np.int64)
    ptr[1:] = np.cumsum([s["host"].shape[0] for s in sel])
    owner = np.array([s["thread"] for s in sel], dtype=np.int32)
    for i, s in enumerate(sel):
        params[ptr[i]:ptr[i + 1], 5] += bases[s["thread"]]
    flop = float(np.sum(2.0 * params[:, 0] * params[:, 1] * params[:, 2]))
    c = np.zeros(base)
    dg = orc.dgemm_ptr()
    t0 = time.perf_counter()
    orc.lib().orc_host_stacks_threaded(params.reshape(-1), ptr, owner, len(sel), cores, A.data, B.data, c, dg)
    dt = time.perf_counter() - t0
    info = {"value": flop / dt * 1e-9, "unit": "GFLOP/s", "cores": cores, "kind": "port",
            "sample": "%d of %d stack entries (first stacks of each of %d threads), %s, stack build %.2fs not included"
                      % (params.shape[0], total_entries, cores, "OpenBLAS dgemm (scipy)" if dg else "naive triple loop", t_build),
            "host_cpus": ncores, "cpu_model": cpu_model()}
    return info


def run_reference(args):
    from dbcsr_b200 import workload

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload.make_config(args.config, nblk=args.nblk)
    vals, ms = [], []
    info = None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        info = cpu_reference_sample(w, args.ref_entries, n_stacks=3 if len(w["sizes"]) <= 3 else len(w["sizes"]))
        if i >= args.warmup:
            vals.append(info["value"])
            ms.append((time.perf_counter() - t0) * 1e3)
    v = float(np.mean(vals))
    info["value"] = v
    sample_flop = None
    print(json.dumps({"impl": "reference", "metric": METRIC_NAMES[args.config], "value": v, "unit": "GFLOP/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "strong",
                      "vs_baseline": None, "dtype": "f64", "data": "synthetic (seed 42)",
                      "config": workload_config(w), "cpu_baseline": info,
                      "e2e": {"value": v, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(w, extra=None):
    c = {"workload": "%s: %dx%d block grid, blocks %s %s, %.0f%% occupation, C=A*B (alpha=1, beta=0)"
                     % (w["name"], w["nblk"], w["nblk"], "x".join(str(s) for s in w["sizes"]) if len(w["sizes"]) > 1 else "%dx%d" % (w["sizes"][0], w["sizes"][0]),
                        "BF16 (FP32 accumulate)" if w["name"] == "cfg4" else "FP64", 100 * w["occupation"]),
         "a_blocks": w["A"].nblks, "b_blocks": w["B"].nblks, "mm_stack_size": 30000,
         "l2_policy": "inputs (A+B %.0f MB, C rewritten every step) larger than L2; C memset each step" % ((w["A"].data.nbytes + w["B"].data.nbytes) / 1e6)}
    if extra:
        c.update(extra)
    return c


# ------------------------------------------------------------------------------------------------ checker (outside timed regions)
def probe_c_blocks(A, B, coords, got_blocks, n_probe=1000, seed=7):
    """Element-wise check of randomly probed C blocks against the oracle's block product (orc_block_gemm: plain loops, the
    arithmetic of the reference's CPU driver).  coords = [(global_row, global_col)], got_blocks(i) -> the device result of block i
    as a flat column-major array.  Returns (number probed, worst relative Frobenius error of a block).  Checker only."""
    from oracle import oracle as orc

    L = orc.lib()
    n = len(coords)
    if n == 0:
        return 0, 0.0
    rng = np.random.default_rng(seed)
    pick = rng.choice(n, size=min(n_probe, n), replace=False)
    a_row_start = np.searchsorted(A.rows, np.arange(1, A.row_sizes.size + 2))      # BCSR: rows ascending
    b_order = np.argsort(B.cols, kind="stable")                                     # B blocks grouped by column, rows ascending
    b_col_start = np.searchsorted(B.cols[b_order], np.arange(1, B.col_sizes.size + 2))
    worst = 0.0
    for i in pick:
        r, c = coords[int(i)]
        m, nn = int(A.row_sizes[r - 1]), int(B.col_sizes[c - 1])
        exp = np.zeros(m * nn)
        ia = np.arange(a_row_start[r - 1], a_row_start[r])
        ib = b_order[b_col_start[c - 1]:b_col_start[c]]
        common, xa, xb = np.intersect1d(A.cols[ia], B.rows[ib], return_indices=True)
        for kk, qa, qb in zip(common, ia[xa], ib[xb]):
            k = int(A.col_sizes[kk - 1])
            L.orc_block_gemm(m, nn, k, A.data[A.offsets[qa]:A.offsets[qa] + m * k], 0, B.data[B.offsets[qb]:B.offsets[qb] + k * nn], exp)
        got = np.asarray(got_blocks(int(i)), dtype=np.float64).reshape(-1)
        den = float(np.linalg.norm(exp))
        worst = max(worst, float(np.linalg.norm(got - exp)) / max(den, 1e-300))
    return int(pick.size), worst


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def gpu_baseline_reference_kernels(bsz=23, nblk=1000, occ=0.1, timeout=90):
    """Same-box GPU baseline: the reference's OWN CUDA backend (libsmm_acc, NVRTC-JIT kernels, H100 parameter set, built for
    compute_100 by baseline/Makefile into baseline/_ref) draining the same kind of stacks through the same ABI, driven by
    tools/kbench (KBENCH_ACC_LIB).  Kernel-only, like `roofline.kernel_only_gflops`."""
    ref = os.path.join(ROOT, "baseline", "_ref", "libdbcsr_acc_ref.so")
    kb = os.path.join(ROOT, "tools", "kbench")
    ours = os.path.join(ROOT, "dbcsr_b200", "lib", "libdbcsr_acc_b200.so")
    if not (os.path.exists(ref) and os.path.exists(kb)):
        return {"unavailable": "baseline/_ref/libdbcsr_acc_ref.so or tools/kbench not built (python -c 'import __graft_entry__ as g; g.build()')"}
    env = dict(os.environ, KBENCH_ACC_LIB=ref)
    try:
        out = subprocess.run([kb, ours, os.path.join(ROOT, "gpurun_out"), str(nblk), str(occ), "3", str(bsz), "0:0:0"], env=env, capture_output=True,
                             text=True, timeout=timeout)
    except Exception as ex:
        return {"unavailable": repr(ex)[:200]}
    import re

    m = re.search(r"parity (\S+).*?mean ([\d.]+) ms.*?([\d.]+) TFLOP/s", out.stdout)
    host_ok = out.stdout.count(": exact")
    if out.returncode != 0 or not m:
        return {"unavailable": "kbench rc %d: %s" % (out.returncode, (out.stderr or out.stdout)[-300:])}
    return {"value": float(m.group(3)) * 1e3, "unit": "GFLOP/s", "ms_per_multiply": float(m.group(2)), "kind": "reference libsmm_acc kernels (NVRTC, parameters_H100.json) on this GPU, kernel-only",
            "workload": "%dx%d block grid, %d^3 blocks, %.0f%% occupation, 30000-entry C-sorted stacks (tools/kbench)" % (nblk, nblk, bsz, 100 * occ),
            "host_checked_blocks_exact": host_ok}


# ------------------------------------------------------------------------------------------------ our arm, one GPU
class Fp64Run:
    """One FP64 config on one GPU: stacks pre-built (one host thread = the reference's traversal order) and resident, panels
    resident, drained through libsmm_acc_process."""

    def __init__(self, acc, cfg_name, nblk, s, nstreams=1, dev_tile=0):
        from dbcsr_b200 import host, workload

        self.acc, self.s = acc, s
        # further streams for the drain (stack i runs on stream i mod nstreams): DBCSR drives one stream per OpenMP thread, so stacks
        # of different threads overlap on the device; 1 = everything on the bench stream (the headline configuration)
        self.side = [acc.stream_create("bench side %d" % i, 0) for i in range(1, nstreams)]
        for st_ in self.side:
            acc.stream_chain(st_, True)
        self.ev_fork = acc.event_create()
        self.ev_join = [acc.event_create() for _ in self.side]
        self.w = w = workload.make_config(cfg_name, nblk=nblk)
        A, B, bs = w["A"], w["B"], w["m_sizes"]
        self.n_st = 3 if len(w["sizes"]) <= 3 else len(w["sizes"])
        self.d_a = acc.to_device(A.data, s)
        self.d_b = acc.to_device(B.data, s)
        host.transpose_panel(acc, B.list3(), bs, bs, self.d_b.ptr, s)
        self.dev_tile = dev_tile
        if dev_tile:
            # stacks of the DEVICE-side builder in tile order (include/dbcsr_b200_host.h, dev_tile): same C index and products as the
            # reference's traversal, ordered by dev_tile x dev_tile squares of C blocks; recorded from one real multiply on the device
            acc.stream_sync(s)
            eng = host.Engine(bs, bs, bs, nthreads=1, mode=host.LAUNCH | host.RECORD | host.DEVICE_BUILD,
                              cfg=host.default_cfg(n_stacks=self.n_st, dev_tile=dev_tile))
            eng.multiply(A.list3(), self.d_a.ptr, B.list3(), self.d_b.ptr)
            eng.sync()
            assert eng.device_built_ticks == 1
        else:
            eng = host.Engine(bs, bs, bs, nthreads=1, mode=host.RECORD, cfg=host.default_cfg(n_stacks=self.n_st))
            eng.multiply(A.list3(), None, B.list3(), None)
        self.stacks = eng.stacks()
        self.flop = eng.flop()
        self.c_rows, self.c_cols, self.c_blk_p, self.c_datasize = eng.c_index(0)
        self.c_rows, self.c_cols, self.c_blk_p = self.c_rows.copy(), self.c_cols.copy(), self.c_blk_p.copy()
        eng.close()
        self.n_entries = sum(st["dev"].shape[0] for st in self.stacks)
        all_dev = np.concatenate([st["dev"].reshape(-1) for st in self.stacks]).astype(np.int32)
        self.d_st = acc.to_device(all_dev, s)
        self.offs = np.concatenate([[0], np.cumsum([st["dev"].size for st in self.stacks])]).astype(np.int64)
        self.alg_bytes = algorithmic_bytes(self.stacks)
        self.runs = sum(1 + int(np.count_nonzero(st["dev"][1:, 2] != st["dev"][:-1, 2])) for st in self.stacks if st["dev"].shape[0])
        # Two pooled C buffers: while the stacks of step k accumulate into buffer k%2, the buffer of step k+1 is zeroed on a side
        # stream (DBCSR zeroes its pooled device C buffer asynchronously at accdrv_init, src/mm/dbcsr_mm_accdrv.F:209-216).  Every
        # step still contains exactly one full memset and waits for it before it ends.
        nbytes = 8 * max(self.c_datasize, 1)
        self.d_cs = [acc.dev_alloc(nbytes), acc.dev_alloc(nbytes)]
        self.zs = acc.stream_create("bench zero", 0)
        self.ev_zero = [acc.event_create(), acc.event_create()]
        self.ev_free = [acc.event_create(), acc.event_create()]
        acc.memset_zero(self.d_cs[0], self.zs)
        acc.event_record(self.ev_zero[0], self.zs)
        acc.event_record(self.ev_free[1], s)
        self.step_no = 0
        self.trickle_ctas = 8

    def drain(self, d_c):
        acc, s = self.acc, self.s
        streams = [s] + self.side
        if self.side:  # fork: the side streams start behind everything enqueued on the bench stream so far
            acc.event_record(self.ev_fork, s)
            for st_ in self.side:
                acc.stream_wait_event(st_, self.ev_fork)
        for i, st in enumerate(self.stacks):
            q = streams[i % len(streams)]
            rc = acc.process(None, self.d_st.ptr + 4 * int(self.offs[i]), st["dev"].shape[0], self.d_a.ptr, self.d_b.ptr, d_c.ptr, st["max_m"],
                             st["max_n"], st["max_k"], st["defined_mnk"], q, q)
            if rc < 0:
                raise RuntimeError("libsmm_acc_process returned %d for stack %d" % (rc, i))
        for st_, ev in zip(self.side, self.ev_join):  # join
            acc.event_record(ev, st_)
            acc.stream_wait_event(s, ev)

    def one_step_trickle(self):
        self.one_step(trickle=True)

    def one_step(self, trickle=False):
        acc, s, zs = self.acc, self.s, self.zs
        k = self.step_no % 2
        self.step_no += 1
        acc.stream_wait_event(zs, self.ev_free[1 - k])     # the other buffer's last reader (step k-1) has finished
        if trickle:  # the same zeros written by a handful of CTAs over the length of the drain instead of one burst
            acc.memset_zero_trickle(self.d_cs[1 - k], zs, nctas=self.trickle_ctas)
        else:
            acc.memset_zero(self.d_cs[1 - k], zs)
        acc.event_record(self.ev_zero[1 - k], zs)
        acc.stream_wait_event(s, self.ev_zero[k])          # zeroed during the previous step
        self.drain(self.d_cs[k])
        acc.event_record(self.ev_free[k], s)
        acc.stream_wait_event(s, self.ev_zero[1 - k])      # the step owns the memset it issued

    def one_step_serial(self):
        """Same work as one_step with the memset IN LINE: zero the C buffer, then drain into it, all on the bench stream.  The
        overlapped variant hides the memset behind the drain but makes both fight for HBM; which one is faster is measured."""
        self.acc.memset_zero(self.d_cs[0], self.s)
        self.drain(self.d_cs[0])

    def restore_overlap_state(self):
        """Leave the double-buffering state as one_step expects it: the buffer of the next step zeroed, nothing in flight."""
        acc = self.acc
        acc.stream_sync(self.s)
        acc.memset_zero(self.d_cs[self.step_no % 2], self.zs)
        acc.event_record(self.ev_zero[self.step_no % 2], self.zs)
        acc.event_record(self.ev_free[1 - self.step_no % 2], self.s)
        acc.stream_sync(self.zs)
        acc.stream_sync(self.s)

    def selfcheck(self, n_probe):
        """Full size, outside every timed region: one drain into a zeroed buffer; (1) sum(C) = colsum(A) . rowsum(B),
        (2) n_probe random C blocks element-wise against the oracle's block product."""
        from dbcsr_b200.cannon import _axis_sums

        acc, s, w = self.acc, self.s, self.w
        A, B = w["A"], w["B"]
        acc.stream_wait_event(s, self.ev_zero[self.step_no % 2])
        acc.stream_sync(self.zs)
        acc.memset_zero(self.d_cs[0], s)
        self.drain(self.d_cs[0])
        c = acc.to_host(self.d_cs[0], (max(self.c_datasize, 1),), np.float64, s)
        got = float(c[:self.c_datasize].sum())
        exp = float(np.dot(_axis_sums(A, 0, A.row_sizes.size, 0), _axis_sums(B, 0, B.col_sizes.size, 1)))
        out = {"property": "sum(C) == colsum(A) . rowsum(B)", "rel_err": abs(got - exp) / max(abs(exp), 1e-300)}
        bs = w["m_sizes"]
        coords = list(zip(self.c_rows.tolist(), self.c_cols.tolist()))

        def got_block(i):
            o = int(self.c_blk_p[i]) - 1
            return c[o:o + int(bs[self.c_rows[i] - 1]) * int(bs[self.c_cols[i] - 1])]

        npr, worst = probe_c_blocks(A, B, coords, got_block, n_probe=n_probe)
        out.update({"probed_blocks": npr, "probe_max_rel_err": worst, "probe": "random C blocks, element-wise vs oracle orc_block_gemm (tolerance 1e-10)",
                    "ok": bool(out["rel_err"] <= 1e-9 and worst <= 1e-10)})
        self.restore_overlap_state()
        return out

    def free_c(self):
        for d in self.d_cs:
            d.free()
        self.d_cs = []

    def close(self):
        self.free_c()
        for d in (self.d_a, self.d_b, self.d_st):
            d.free()
        self.acc.stream_destroy(self.zs)
        for st_ in self.side:
            self.acc.stream_chain(st_, False)
            self.acc.stream_destroy(st_)


def timed_steps(torch, tstream, acc, s, fn, steps, lookahead=2):
    """CUDA-event time of `steps` calls of fn() on the bench stream.  The host enqueues at most `lookahead` steps ahead of the device
    (it waits for the end event of step k - lookahead before enqueuing step k): the device never idles between steps, and the launch
    queue never holds more than a few hundred kernels."""
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(steps)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(tstream):
        for k in range(steps):
            if lookahead and k >= lookahead:
                ev[k - lookahead][1].synchronize()
            ev[k][0].record(tstream)
            fn()
            ev[k][1].record(tstream)
    acc.stream_sync(s)
    torch.cuda.synchronize()
    return [ev[k][0].elapsed_time(ev[k][1]) for k in range(steps)], time.perf_counter() - t0


def measure_fp64_peaks(torch, acc, s):
    """Denominators measured in this run on this GPU: the DMMA.8x8x4 register-operand loop of the library (what the stack kernel's
    pipe can do) and cuBLAS DGEMM 8192^3 through torch.matmul (what NVIDIA's own FP64 GEMM reaches)."""
    dmma = acc.fp64_peak_gflops(s)
    dmma_sustained = acc.fp64_peak_sustained_gflops(s, 0.4)

    dgemm = None
    dgemm_sustained = None
    try:
        n = 8192
        x = torch.rand((n, n), dtype=torch.float64, device="cuda")
        y = torch.rand((n, n), dtype=torch.float64, device="cuda")
        best = 1e30
        for i in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(x, y)
            e1.record()
            e1.synchronize()
            if i:
                best = min(best, e0.elapsed_time(e1))
        dgemm = 2.0 * n ** 3 / (best * 1e-3) * 1e-9
        # back to back for ~0.4 s on the same uniform(0,1) data the bench feeds the stack kernel: what cuBLAS sustains under the power limit
        reps = max(4, int(0.4 / (best * 1e-3)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(reps):
            if i == reps // 2:
                e0.record()
            torch.matmul(x, y)
        e1.record()
        e1.synchronize()
        dgemm_sustained = 2.0 * n ** 3 * (reps - reps // 2) / (e0.elapsed_time(e1) * 1e-3) * 1e-9
        del x, y
        torch.cuda.empty_cache()
    except Exception as ex:
        print("bench: cuBLAS DGEMM peak probe failed: %r" % (ex,), file=sys.stderr)
    return {"dmma_burst": dmma, "dmma_sustained": dmma_sustained, "dgemm": dgemm, "dgemm_sustained": dgemm_sustained}


def fp64_config_report(torch, tstream, acc, s, run, steps, warmup, n_probe, peaks, ncu_key):
    """Warm-up, self-check, timed steps (value) and kernel-only drains (roofline) of one FP64 config."""
    for _ in range(warmup):
        run.one_step()
    acc.stream_sync(s)
    check = run.selfcheck(n_probe) if n_probe else None
    # how the per-step memset of C is scheduled is the caller's choice (DBCSR zeroes its pooled buffer asynchronously): measure both
    # the overlapped and the in-line variant on three steps each and time the faster one
    # (interleaved A B A B after the warm-up, so that both see the same clocks: the board's power limit pulls the SM clock down within
    # ~50 ms of sustained load)
    modes = {"overlapped on a side stream (cudaMemsetAsync)": run.one_step, "overlapped on a side stream at a bounded rate (8 CTAs)": run.one_step_trickle,
             "in line on the bench stream": run.one_step_serial}
    trials = {name: [] for name in modes}
    for _ in range(2):
        for name, fn in modes.items():
            t_, _ = timed_steps(torch, tstream, acc, s, fn, 3)
            trials[name] += t_
            run.restore_overlap_state()
    trial_ms = {name: float(np.mean(v)) for name, v in trials.items()}
    zero_mode = min(trial_ms, key=trial_ms.get)
    step_fn = modes[zero_mode]
    run.restore_overlap_state()
    launches0 = acc.launch_count()
    step_ms, t_wall = timed_steps(torch, tstream, acc, s, step_fn, steps)
    launches = acc.launch_count() - launches0
    kern_ms, _ = timed_steps(torch, tstream, acc, s, lambda: run.drain(run.d_cs[0]), 3)
    # the same drain as a burst: one drain after the device has idled for half a second (clocks back at maximum)
    time.sleep(0.5)
    t_enq0 = time.perf_counter()
    burst_ms, _ = timed_steps(torch, tstream, acc, s, lambda: run.drain(run.d_cs[0]), 1)
    acc.stream_sync(s)
    t_enq0 = time.perf_counter()
    run.drain(run.d_cs[0])
    host_enqueue_us = (time.perf_counter() - t_enq0) * 1e6 / max(len(run.stacks), 1)  # host time per libsmm_acc_process call (Python + ctypes + launch)
    acc.stream_sync(s)
    # how the drain time develops under continuous load: 12 drains back to back after another idle half second
    time.sleep(0.5)
    series_ms, _ = timed_steps(torch, tstream, acc, s, lambda: run.drain(run.d_cs[0]), 12)
    ms_per_step = float(np.mean(step_ms))
    k_ms = float(np.mean(kern_ms))
    value = run.flop / (ms_per_step * 1e-3) * 1e-9
    kernel_only = run.flop / (k_ms * 1e-3) * 1e-9
    hbm_peak, hbm_src = measured_peaks()
    dmma_peak, dmma_burst, dgemm_peak = peaks["dmma_sustained"], peaks["dmma_burst"], peaks["dgemm"]
    kernel_burst = run.flop / (float(burst_ms[0]) * 1e-3) * 1e-9
    nst = max(len(run.stacks), 1)
    traffic = None
    ncu_json = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(ncu_json):
        try:
            traffic = json.load(open(ncu_json)).get(ncu_key, {}).get("dram_bytes_per_launch") if ncu_key else None
        except Exception:
            traffic = None
    st0 = max(run.stacks, key=lambda st: st["dev"].shape[0] * st["m"] * st["n"] * st["k"])
    m0, n0, k0 = st0["m"], st0["n"], st0["k"]
    pad = (m0 * n0 * k0) / float(((m0 + 7) // 8 * 8) * ((n0 + 7) // 8 * 8) * ((k0 + 3) // 4 * 4))
    launch_us = k_ms * 1e3 / nst
    alg_gbs = run.alg_bytes / (k_ms * 1e-3) * 1e-9
    roofline = {
        "bound": "tensor", "pipe": "FP64 tensor pipe (DMMA.8x8x4 = mma.sync.m8n8k4.f64; tcgen05 has no f64 kind)",
        "achieved": kernel_only * 1e-3, "peak": dmma_peak * 1e-3, "unit": "TFLOP/s", "frac": kernel_only / dmma_peak if dmma_peak > 0 else None,
        "traffic": traffic,
        "peak_source": "measured in this run: register-operand DMMA loop of the library run back to back for 0.4 s (libsmm_acc_b200_fp64_peak_sustained_gflops) -- the kernel is timed inside long steps, under the board's power limit; burst figures and cuBLAS DGEMM 8192^3 beside it",
        "burst": {"note": "one drain / one peak-probe launch after the device idled (SM clock at maximum)", "kernel_only_gflops": kernel_burst,
                  "peak_gflops": dmma_burst, "frac": kernel_burst / dmma_burst if dmma_burst > 0 else None},
        "cublas_dgemm_8192_gflops": dgemm_peak,
        "cublas_dgemm_8192_sustained_gflops": peaks.get("dgemm_sustained"),
        "kernel": "smm_dmma_kernel<%d,%d,%d> (dominant of %d launches/step)" % (m0, n0, k0, nst),
        "algorithmic_flop_per_launch": run.flop / nst, "avg_launch_us": launch_us, "kernel_only_gflops": kernel_only,
        "host_enqueue_us_per_launch": host_enqueue_us,
        "drain_series_after_idle_ms": [round(float(x), 3) for x in series_ms],
        "frac_on_timed_value": value / dmma_peak if dmma_peak > 0 else None,
        "alt_bounds": {
            "padded_tensor_ceiling": {"note": "tiles of 8x8x4: %dx%dx%d is %.0f%% useful" % (m0, n0, k0, 100 * pad), "frac": kernel_only / (dmma_peak * pad) if dmma_peak > 0 else None},
            "hbm_streaming_model": {"note": "SURVEY 8(d) bytes: A+B block and 12 B per entry, C read+write per run of equal c_first; frac > 1 means operands are re-used out of L2",
                                    "achieved": alg_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": alg_gbs / hbm_peak, "peak_source": hbm_src,
                                    "algorithmic_bytes_per_launch": run.alg_bytes / nst, "mean_run_length": run.n_entries / max(1, run.runs)},
            "hbm_real_traffic": ({"note": "dram__bytes_read+write per launch of the shipped kernel (ncu --set full, profiles/ncu_summary.json)",
                                  "achieved": traffic / (launch_us * 1e-6) * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                                  "frac": traffic / (launch_us * 1e-6) * 1e-9 / hbm_peak} if traffic else None)}}
    return {"value": value, "ms_per_step": ms_per_step, "kernel_only_gflops": kernel_only, "launches": int(launches), "wall_s": t_wall, "roofline": roofline,
            "zero_mode": zero_mode, "zero_mode_trial_ms": trial_ms,
            "selfcheck": check, "products": run.n_entries, "flop": run.flop, "stacks": len(run.stacks), "c_blocks": int(run.c_rows.size)}


def bf16_config_report(torch, acc, s, tstream, nblk, steps, warmup, n_probe=200):
    """BASELINE config 4 (23x23 blocks, 50 % occupation, BF16 operands / FP32 C) through the tiled tcgen05 SpGEMM
    (libsmm_acc_b200_bf16_spgemm): one launch per multiply, C tiles accumulate in TMEM and are written once."""
    from dbcsr_b200 import workload
    from dbcsr_b200.bf16 import Bf16SpGemm

    w = workload.make_config("cfg4", nblk=nblk)
    A, B = w["A"], w["B"]
    mm = Bf16SpGemm(acc, A, B, s)
    m, n = mm.m, mm.n
    for _ in range(warmup):
        mm.run()
    acc.stream_sync(s)
    # parity (outside the timed region): probed C blocks against the FP64 oracle on the UNROUNDED inputs, tolerance 1e-3
    check = None
    if n_probe:
        c = mm.result()
        coords = [(r + 1, cc + 1) for r in range(mm.nrb) for cc in range(mm.ncb)]
        npr, worst = probe_c_blocks(A, B, coords, lambda i: c[i // mm.ncb, i % mm.ncb].reshape(-1), n_probe=n_probe)
        check = {"probed_blocks": npr, "probe_max_rel_err": worst, "probe": "random C blocks vs FP64 oracle orc_block_gemm on unrounded inputs (tolerance 1e-3)",
                 "ok": bool(worst <= 1e-3)}
        del c
    launches0 = acc.launch_count()
    step_ms, _ = timed_steps(torch, tstream, acc, s, mm.run, steps)
    launches = acc.launch_count() - launches0
    ms = float(np.mean(step_ms))
    value = mm.flop / (ms * 1e-3) * 1e-9
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        tpeak, tsrc = float(pk["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained: cuBLAS bf16 8192^3 back to back)"
    except Exception:
        tpeak, tsrc = 1590.0, "fallback (B200_PROFILING.md)"
    hbm_peak, _ = measured_peaks()
    tile_bytes = acc.bf16_rk_tile_bytes(m)
    compulsory = (A.nblks + B.nblks) * tile_bytes + 4 * mm.c_elems  # every operand tile read once, C written once
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json"))).get("cfg4_planned", {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "tensor", "pipe": "tcgen05.mma kind::f16 (BF16 operands, FP32 accumulators in TMEM)", "achieved": value * 1e-3, "peak": tpeak,
                "unit": "TFLOP/s", "frac": value * 1e-3 / tpeak, "traffic": traffic, "peak_source": tsrc,
                "kernel": ("bt_plan_kernel + smm_bf16_planned_kernel (2 launches/step: copy commands and MMA runs per (group, k block) derived once, then the "
                           "multiply; M=128 x N=32r x K=16 MMAs, 2 per run of r adjacent existing B blocks)") if acc.get_tunable("bf16_plan") else
                          "smm_bf16_tiled_kernel (1 launch/step; M=128 x N=32r x K=16 MMAs, 2 per run of r adjacent existing B blocks, tile and k block)",
                "useful_flop_per_launch": mm.flop, "issued_flop_per_launch": mm.issued_flop, "issued_frac": mm.issued_flop / (ms * 1e-3) * 1e-12 / tpeak,
                "useful_over_issued": mm.flop / max(mm.issued_flop, 1),
                "alt_bounds": {"hbm_compulsory": {"note": "operand tiles once + C once", "bytes": compulsory, "achieved": compulsory / (ms * 1e-3) * 1e-9,
                                                  "peak": hbm_peak, "unit": "GB/s", "frac": compulsory / (ms * 1e-3) * 1e-9 / hbm_peak}}}
    out = {"metric": METRIC_NAMES["cfg4"], "value": value, "unit": "GFLOP/s", "ms_per_step": ms, "dtype": "bf16", "kernel_only_gflops": value,
           "config": workload_config(w, {"products": mm.products, "flop": mm.flop, "timed": "CUDA events around one libsmm_acc_b200_bf16_spgemm call per step (plan kernel + multiply kernel; C is overwritten, no memset)"}),
           "gpu_launches": int(launches), "roofline": roofline, "selfcheck": check}
    mm.close()
    return out


def run_single_bf16(args):
    import torch

    from dbcsr_b200 import lib as acclib

    acc = acclib.Acc(0)
    s = acc.stream_create("bench", 0)
    tstream = torch.cuda.ExternalStream(acclib.ctypes.c_void_p.from_address(s).value)
    sampler = ClockSampler(0)
    sampler.start()
    rep = bf16_config_report(torch, acc, s, tstream, args.nblk, args.steps, args.warmup, 0 if args.no_selfcheck else 200)
    clocks = sampler.stop()
    rep.update({"n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "data": "synthetic (numpy PCG64 seed 42, uniform(0,1))", "clocks": clocks, "e2e": None, "cpu_baseline": None})
    print(json.dumps(rep))
    acc.stream_destroy(s)


def run_single(args):
    import torch

    from dbcsr_b200 import host
    from dbcsr_b200 import lib as acclib
    from dbcsr_b200.multiply import DeviceMultiply

    if args.config == "cfg4":
        return run_single_bf16(args)
    acc = acclib.Acc(0)
    s = acc.stream_create("bench", 0)
    # the bench stream carries nothing but stack drains between event waits: declare it a chain (programmatic dependent launch
    # without the grid-dependency wait in front of the reads, include/dbcsr_acc_libsmm.h)
    if not args.no_chain:
        acc.stream_chain(s, True)
    tstream = torch.cuda.ExternalStream(acclib.ctypes.c_void_p.from_address(s).value)
    if args.no_peak_probes:  # launch-list runs under ncu: keep the probes' kernels out of the list (the line's roofline is then NOT measured)
        peaks = {"dmma_burst": 36900.0, "dmma_sustained": 36900.0, "dgemm": None, "dgemm_sustained": None, "not_measured": True}
    else:
        peaks = measure_fp64_peaks(torch, acc, s)
    run = Fp64Run(acc, args.config, args.nblk, s, nstreams=(args.cfg3_streams if args.config == "cfg3" else 1))
    w = run.w
    A, B, bs = w["A"], w["B"], w["m_sizes"]

    sampler = ClockSampler(0)
    if not args.no_clock_sampler:
        sampler.start()
    time.sleep(0.3)
    rep = fp64_config_report(torch, tstream, acc, s, run, args.steps, args.warmup, 0 if args.no_selfcheck else args.probe_blocks, peaks, args.config)
    clocks = sampler.stop()

    # ---- end to end through the host engine, host buffers pinned
    e2e = None
    if not args.no_e2e:
        try:
            run.free_c()
            pa = acc.host_alloc((A.data.size,), np.float64)
            pb = acc.host_alloc((B.data.size,), np.float64)
            pa.array[:] = A.data
            pb.array[:] = B.data
            a_l, b_l = A.list3(), B.list3()

            def e2e_leg(builder):
                """One variant of the end-to-end multiply: `host` = multi-threaded host stack builder (stacks uploaded),
                `device` = device-side builder (index lists uploaded, stacks never leave the device)."""
                if builder == "device":
                    nthreads, rchunks, mode = args.dev_threads, args.dev_row_chunks, host.LAUNCH | host.DEVICE_BUILD
                else:
                    nthreads, rchunks, mode = args.threads or min(32, max(1, (os.cpu_count() or 2) // 2)), args.row_chunks, host.LAUNCH
                cfg_e2e = host.default_cfg(n_stacks=run.n_st, row_chunks=rchunks, dev_tile=(args.dev_tile if builder == "device" else 0))
                # left panel uploaded in block-row pieces behind the right one: measured neutral with the host builder and SLOWER with
                # the device builder (101.4 vs 96.9 ms: the per-chunk upload bookkeeping costs more than the overlap gains) -- opt-in
                pipelined = bool(args.pipelined_upload)
                dm = DeviceMultiply(acc, bs, bs, bs, A.data.size, B.data.size, B.nblks, nthreads=nthreads, cfg=cfg_e2e, mode=mode)
                pcs = None
                times = []
                try:
                    for it in range(max(1, args.e2e_warmup) + args.e2e_steps):
                        acc.device_synchronize()
                        t0 = time.perf_counter()
                        dm.upload_panels(pa.array, pb.array, b_l, a_list3=a_l if pipelined else None)
                        t_up = time.perf_counter()
                        dm.multiply(a_l, b_l)
                        t_mul = time.perf_counter()
                        if pcs is None:  # pinned result buffers, sized at the first pass (pooled afterwards, like DBCSR's memory pools);
                            # ONE pinned pool serves both variants: pinning 4 GB is the slowest host operation of the whole bench
                            dm.engine.sync()
                            caps = [max(dm.engine.c_capacity(t), 1) for t in range(nthreads)]
                            if pool[0] is None or pool[0].array.size < sum(caps):
                                if pool[0] is not None:
                                    pool[0].free()
                                pool[0] = acc.host_alloc((sum(caps),), np.float64)
                            offs = np.concatenate([[0], np.cumsum(caps)])
                            pcs = [pool[0].array[int(offs[t]):int(offs[t + 1])] for t in range(nthreads)]
                            prod = dm.download_c(pcs)
                            dm.set_result_buffers(pcs)
                        else:
                            prod = dm.download_c()
                        dt = time.perf_counter() - t0
                        if it >= max(1, args.e2e_warmup):
                            times.append(dt)
                            phases = {"enqueue_upload_ms": (t_up - t0) * 1e3, "host_build_and_enqueue_ms": (t_mul - t_up) * 1e3,
                                      "drain_and_d2h_ms": (time.perf_counter() - t_mul) * 1e3}
                    # the product of the last step, checked where it arrived (pinned host buffers): sum property
                    got = sum(float(p[3].sum()) for p in prod.parts)
                    stack_bytes = 12 * run.n_entries if builder == "host" else 12 * (A.nblks + B.nblks)
                    leg = {"value": run.flop / float(np.mean(times)) * 1e-9, "unit": "GFLOP/s", "h2d_bytes_per_step": int(dm.h2d_bytes + stack_bytes),
                           "d2h_bytes_per_step": int(dm.d2h_bytes), "ms_per_step": float(np.mean(times)) * 1e3, "stack_builder": builder,
                           "host_threads": nthreads, "row_chunks_per_thread": rchunks, "pipelined_upload": pipelined,
                           "host_build_seconds": dm.engine.build_seconds(), "c_blocks": prod.nblks, "phases_last_step": phases,
                           "device_built_ticks": dm.engine.device_built_ticks, "sum_c": got, "dev_tile": (args.dev_tile if builder == "device" else 0),
                           "timing": "wall clock around the public call, device synchronised on both sides"}
                finally:
                    dm.close()
                return leg

            pool = [None]
            legs = {}
            for b_ in (["host", "device"] if args.e2e_builder == "both" else [args.e2e_builder]):
                try:
                    legs[b_] = e2e_leg(b_)
                except Exception as ex:
                    import traceback

                    traceback.print_exc(file=sys.stderr)
                    legs[b_] = {"value": None, "error": repr(ex)[:300], "stack_builder": b_}
            ok = [v for v in legs.values() if v.get("value")]
            if not ok:
                raise RuntimeError("no e2e variant ran: %r" % ({k: v.get("error") for k, v in legs.items()},))
            if len(ok) == 2 and abs(ok[0]["sum_c"] / ok[1]["sum_c"] - 1.0) > 1e-12:
                raise RuntimeError("e2e variants disagree: sum(C) %r vs %r" % (ok[0]["sum_c"], ok[1]["sum_c"]))
            e2e = dict(max(ok, key=lambda v: v["value"]))
            e2e["variants"] = {k: {kk: v.get(kk) for kk in ("value", "ms_per_step", "host_threads", "row_chunks_per_thread", "host_build_seconds",
                                                          "phases_last_step", "h2d_bytes_per_step", "error") if v.get(kk) is not None}
                               for k, v in legs.items()}
            for p_ in [pa, pb, pool[0]]:
                if p_ is not None:
                    p_.free()
        except Exception as ex:  # the headline line must still be printed (e.g. not enough pinned memory on this host) -- but loudly
            import traceback

            traceback.print_exc(file=sys.stderr)
            print("bench: the end-to-end leg FAILED: %r" % (ex,), file=sys.stderr)
            e2e = {"value": None, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": repr(ex)[:300]}

    cpu = None
    if not args.no_cpu:
        try:
            cpu = cpu_reference_sample(w, args.ref_entries, n_stacks=run.n_st)
        except Exception as ex:
            cpu = {"value": None, "unit": "GFLOP/s", "cores": 0, "kind": "port", "sample": "failed: " + repr(ex)[:200]}
    run.close()

    # ---- the same multiply on the device builder's TILE-ORDERED stacks (not the reference's stack contents: same products and C index,
    #      every C block accumulated in one run, operands L2 resident per square of C blocks) -- what an engine that builds its stacks
    #      on the device can feed the same kernel; reported beside the headline, never instead of it
    tiled = None
    if args.config == "cfg2" and args.dev_tile > 0 and not args.no_tiled:
        try:
            rt = Fp64Run(acc, args.config, args.nblk, s, dev_tile=args.dev_tile)
            rept = fp64_config_report(torch, tstream, acc, s, rt, max(3, args.steps // 2), 3, 0 if args.no_selfcheck else args.probe_blocks, peaks, None)
            tiled = {"value": rept["value"], "unit": "GFLOP/s", "ms_per_step": rept["ms_per_step"], "kernel_only_gflops": rept["kernel_only_gflops"],
                     "burst_kernel_only_gflops": rept["roofline"]["burst"]["kernel_only_gflops"], "frac_of_fp64_tensor_peak": rept["roofline"]["frac"],
                     "burst_frac": rept["roofline"]["burst"]["frac"], "dev_tile": args.dev_tile, "stacks": rept["stacks"],
                     "mean_run_length": rt.n_entries / max(1, rt.runs), "selfcheck": rept["selfcheck"], "zero_mode": rept["zero_mode"],
                     "drain_series_after_idle_ms": rept["roofline"]["drain_series_after_idle_ms"],
                     "note": "stacks built by the device-side builder in tile order (cfg.dev_tile): identical C index and products, different "
                             "stack contents than DBCSR's traversal; the headline `value` stays on the reference's stacks"}
            if args.tiled_sweep:  # launch-policy knobs on the tile-ordered stacks: (align, chunk) -> ms per drain, burst and sustained
                sweep = {}
                for al, ch in [(1, 12), (0, 12), (1, 10), (0, 10), (1, 20), (0, 20), (1, 8), (1, 16), (1, 30), (0, 30), (1, 0), (0, 0)]:
                    acc.set_tunable("align", al)
                    acc.set_tunable("chunk", ch)
                    time.sleep(0.3)
                    b_ms, _ = timed_steps(torch, tstream, acc, s, lambda: rt.drain(rt.d_cs[0]), 1)
                    s_ms, _ = timed_steps(torch, tstream, acc, s, lambda: rt.drain(rt.d_cs[0]), 10)
                    sweep["align%d_chunk%d" % (al, ch)] = [round(float(b_ms[0]), 3), round(float(np.mean(s_ms[5:])), 3)]
                acc.set_tunable("align", -1)
                acc.set_tunable("chunk", -1)
                tiled["policy_sweep_ms_burst_sustained"] = sweep
            rt.close()
        except Exception as ex:
            import traceback

            traceback.print_exc(file=sys.stderr)
            tiled = {"error": repr(ex)[:300]}

    # ---- the other single-GPU configs of BASELINE.json at their full size (driver-visible; not bench lines of their own)
    extra = {}
    if args.config == "cfg2" and not args.no_extra and (args.nblk is None or args.extra_nblk is not None):
        try:
            r3 = Fp64Run(acc, "cfg3", args.extra_nblk, s, nstreams=args.cfg3_streams)
            rep3 = fp64_config_report(torch, tstream, acc, s, r3, max(3, args.steps // 2), 3, 0 if args.no_selfcheck else args.probe_blocks, peaks, "cfg3")
            extra["cfg3"] = {"metric": METRIC_NAMES["cfg3"], "value": rep3["value"], "unit": "GFLOP/s", "ms_per_step": rep3["ms_per_step"], "dtype": "f64",
                             "kernel_only_gflops": rep3["kernel_only_gflops"], "config": workload_config(r3.w, {"products": rep3["products"], "flop": rep3["flop"],
                                                                                                           "stacks": rep3["stacks"], "streams": args.cfg3_streams}),
                             "gpu_launches": rep3["launches"], "roofline": rep3["roofline"], "selfcheck": rep3["selfcheck"], "zero_mode": rep3["zero_mode"]}
            r3.close()
        except Exception as ex:
            import traceback

            traceback.print_exc(file=sys.stderr)
            extra["cfg3"] = {"error": repr(ex)[:300]}
        try:
            extra["cfg4"] = bf16_config_report(torch, acc, s, tstream, args.extra_nblk, max(3, args.steps // 2), 3, 0 if args.no_selfcheck else 200)
        except Exception as ex:
            import traceback

            traceback.print_exc(file=sys.stderr)
            extra["cfg4"] = {"error": repr(ex)[:300]}

    gpu_base = None
    if not args.no_gpu_baseline and args.config == "cfg2":
        acc.device_synchronize()
        gpu_base = gpu_baseline_reference_kernels(nblk=args.nblk or 1000)

    out = {"metric": METRIC_NAMES[args.config],
           "value": rep["value"], "unit": "GFLOP/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": rep["ms_per_step"],
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic (numpy PCG64 seed 42, uniform(0,1))",
           "config": workload_config(w, {"products": rep["products"], "flop": rep["flop"], "stacks": rep["stacks"], "c_blocks": rep["c_blocks"],
                                         "timed": "CUDA events on the launching stream; step = one memset of the whole C buffer + %d libsmm_acc_process calls into it; memset %s (trial: %s)" % (rep["stacks"], rep["zero_mode"], json.dumps(rep["zero_mode_trial_ms"])),
                                         "pdl_chain": not args.no_chain}),
           "clocks": clocks, "gpu_launches": rep["launches"], "wall_s_timed_region": rep["wall_s"], "roofline": rep["roofline"], "e2e": e2e, "cpu_baseline": cpu,
           "gpu_baseline": gpu_base, "selfcheck": rep["selfcheck"], "tile_order": tiled, "extra_configs": extra or None}
    print(json.dumps(out))
    acc.stream_destroy(s)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg3", "cfg4"])
    ap.add_argument("--nblk", type=int, default=None, help="override the block-grid size (default 1000)")
    ap.add_argument("--threads", type=int, default=0, help="host threads of the e2e engine (default: min(32, cpus/2))")
    ap.add_argument("--row-chunks", type=int, default=4, help="block-row chunks per host thread in the e2e engine (earlier D2H)")
    ap.add_argument("--pipelined-upload", action="store_true",
                    help="e2e: upload the left panel in block-row chunks behind the right panel (measured neutral on cfg2: 96.2 vs 95.2 ms)")
    ap.add_argument("--e2e-builder", default="both", choices=["host", "device", "both"],
                    help="stack builder of the e2e leg: multi-threaded host builder, device-side builder, or both (the faster one is reported)")
    ap.add_argument("--dev-threads", type=int, default=2, help="host threads (= independent device build pipelines) of the device-builder e2e leg")
    ap.add_argument("--dev-row-chunks", type=int, default=4, help="block-row slices per thread of the device-builder e2e leg (early D2H)")
    ap.add_argument("--dev-tile", type=int, default=64, help="square size (C blocks) of the tile-order leg; 0 = skip")
    ap.add_argument("--no-tiled", action="store_true", help="skip the tile-order leg")
    ap.add_argument("--no-peak-probes", action="store_true", help="do not run the DMMA / cuBLAS peak probes (ncu launch lists); the roofline of the line is then not a measurement")
    ap.add_argument("--tiled-sweep", action="store_true", help="sweep the (align, chunk) launch knobs on the tile-ordered stacks (diagnostic)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-warmup", type=int, default=1)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-selfcheck", action="store_true", help="skip the full-size sum(C) / probed-block check before the timed region")
    ap.add_argument("--probe-blocks", type=int, default=1000, help="random C blocks checked element-wise against the oracle in the self-check")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_configs legs (cfg3, cfg4 at full size)")
    ap.add_argument("--cfg3-streams", type=int, default=4, help="streams the 125 stacks of the cfg3 extra config are spread over (DBCSR: one per OpenMP thread)")
    ap.add_argument("--extra-nblk", type=int, default=None, help="block-grid size of the extra_configs legs (default: the full 1000)")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the reference-kernels-on-this-GPU leg")
    ap.add_argument("--no-clock-sampler", action="store_true", help="do not poll nvidia-smi during the run (diagnostic)")
    ap.add_argument("--no-chain", action="store_true", help="do not declare the bench stream a chain of independent drains (every kernel waits for its predecessor)")
    ap.add_argument("--ref-entries", type=int, default=4_000_000, help="stack entries in the bounded CPU sample")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 or world > 1:
        from dbcsr_b200 import cannon

        return cannon.bench_main(args)
    return run_single(args)


if __name__ == "__main__":
    main()
