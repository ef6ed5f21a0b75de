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
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
        sm, smmax, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smmax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        load = [s for s in sm if s > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(smmax) if smmax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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
    from dbcsr_b200 import host
    from oracle import oracle as orc

    ncores = os.cpu_count() or 1
    cores = min(cores or ncores, orc.lib().orc_max_threads(), 64)
    A, B, bs = w["A"], w["B"], w["m_sizes"]
    t0 = time.perf_counter()
    eng = host.Engine(bs, bs, bs, nthreads=cores, mode=host.RECORD, cfg=host.default_cfg(mm_stack_size=1000, n_stacks=n_stacks))
    eng.multiply(A.list3(), None, B.list3(), None)
    t_build = time.perf_counter() - t0
    stacks = eng.stacks()
    total_entries = sum(s["host"].shape[0] for s in stacks)
    # per-thread C areas laid out one after the other; sample = the first stacks of every thread up to the entry budget
    bases, base = [], 0
    for t in range(cores):
        bases.append(base)
        base += eng.c_index(t)[3]
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
    eng.close()
    info = {"value": flop / dt * 1e-9, "unit": "GFLOP/s", "cores": cores, "kind": "port",
            "sample": "%d of %d stack entries (first stacks of each of %d threads), %s, stack build %.2fs not included"
                      % (params.shape[0], total_entries, cores, "OpenBLAS dgemm (scipy)" if dg else "naive triple loop", t_build),
            "host_cpus": ncores}
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


# ------------------------------------------------------------------------------------------------ our arm, one GPU
def run_single(args):
    import torch

    from dbcsr_b200 import host, workload
    from dbcsr_b200 import lib as acclib
    from dbcsr_b200.multiply import DeviceMultiply

    acc = acclib.Acc(0)
    w = workload.make_config(args.config, nblk=args.nblk)
    A, B, bs = w["A"], w["B"], w["m_sizes"]
    n_st = 3 if len(w["sizes"]) <= 3 else len(w["sizes"])
    cfg = host.default_cfg(n_stacks=n_st)

    # ---- build the stacks once on the host (one thread => the reference's traversal order), keep device-order copies
    eng = host.Engine(bs, bs, bs, nthreads=1, mode=host.RECORD, cfg=cfg)
    eng.multiply(A.list3(), None, B.list3(), None)
    stacks = eng.stacks()
    flop = eng.flop()
    c_datasize = eng.c_index(0)[3]
    c_nblks = eng.c_index(0)[0].size
    eng.close()
    n_entries = sum(s["dev"].shape[0] for s in stacks)

    s = acc.stream_create("bench", 0)
    raw_stream = acclib.ctypes.c_void_p.from_address(s).value
    tstream = torch.cuda.ExternalStream(raw_stream)
    d_a = acc.to_device(A.data, s)
    d_b = acc.to_device(B.data, s)
    host.transpose_panel(acc, B.list3(), bs, bs, d_b.ptr, s)
    all_dev = np.concatenate([st["dev"].reshape(-1) for st in stacks]).astype(np.int32)
    d_st = acc.to_device(all_dev, s)
    offs = np.concatenate([[0], np.cumsum([st["dev"].size for st in stacks])]).astype(np.int64)
    bf16 = args.config == "cfg4"
    dtype_id = acclib.DBCSR_TYPE_BF16_EXT if bf16 else acclib.DBCSR_TYPE_REAL_8
    if bf16:  # pack both panels once into BF16 operand tiles (A: m x k col-major, B: transposed n x k col-major)
        mm_ = int(w["sizes"][0])
        ta_bytes = acc.bf16_tile_bytes(mm_, mm_)
        p_a, p_b = acc.dev_alloc(A.nblks * ta_bytes), acc.dev_alloc(B.nblks * ta_bytes)
        acc.pack_bf16(d_a.ptr, A.nblks, mm_, mm_, 1, mm_, p_a.ptr, s)
        acc.pack_bf16(d_b.ptr, B.nblks, mm_, mm_, 1, mm_, p_b.ptr, s)
        acc.stream_sync(s)
        d_a.free()
        d_b.free()
        d_a, d_b = p_a, p_b
    d_c = acc.dev_alloc((4 if bf16 else 8) * max(c_datasize, 1))
    alg_bytes = algorithmic_bytes(stacks)

    def drain(d_c):
        for i, st in enumerate(stacks):
            rc = acc.process(None, d_st.ptr + 4 * int(offs[i]), st["dev"].shape[0], d_a.ptr, d_b.ptr, d_c.ptr, st["max_m"], st["max_n"],
                             st["max_k"], st["defined_mnk"], s, s, datatype=dtype_id)
            if rc < 0:
                raise RuntimeError("libsmm_acc_process returned %d for stack %d" % (rc, i))

    # Two pooled C buffers: while the stacks of step k accumulate into buffer k%2, the buffer of step k+1 is zeroed on a side
    # stream (DBCSR zeroes its pooled device C buffer asynchronously at accdrv_init, src/mm/dbcsr_mm_accdrv.F:209-216).  Every
    # step still contains exactly one full memset and waits for it before it ends, so step time = max(drain, memset) + epsilon.
    d_cs = [d_c, acc.dev_alloc((4 if bf16 else 8) * max(c_datasize, 1))]
    zs = acc.stream_create("bench zero", 0)
    ev_zero = [acc.event_create(), acc.event_create()]
    ev_free = [acc.event_create(), acc.event_create()]
    acc.memset_zero(d_cs[0], zs)
    acc.event_record(ev_zero[0], zs)
    acc.event_record(ev_free[1], s)
    step_no = [0]

    def one_step():
        k = step_no[0] % 2
        step_no[0] += 1
        acc.stream_wait_event(zs, ev_free[1 - k])     # the other buffer's last reader (step k-1) has finished
        acc.memset_zero(d_cs[1 - k], zs)
        acc.event_record(ev_zero[1 - k], zs)
        acc.stream_wait_event(s, ev_zero[k])          # zeroed during the previous step
        drain(d_cs[k])
        acc.event_record(ev_free[k], s)
        acc.stream_wait_event(s, ev_zero[1 - k])      # the step owns the memset it issued

    for _ in range(args.warmup):
        one_step()
    acc.stream_sync(s)

    # self-check at full size, outside every timed region: one drain into a zeroed buffer must satisfy the size-independent
    # property sum(C) = colsum(A) . rowsum(B) (FP64 path; tests/test_gpu_multiply.py checks the same property)
    selfcheck = None
    if not bf16 and not args.no_selfcheck:
        try:
            from dbcsr_b200.cannon import _axis_sums

            acc.stream_wait_event(s, ev_zero[step_no[0] % 2])
            acc.stream_sync(zs)
            acc.memset_zero(d_cs[0], s)
            drain(d_cs[0])
            got = float(acc.to_host(d_cs[0], (max(c_datasize, 1),), np.float64, s)[:c_datasize].sum())
            exp = float(np.dot(_axis_sums(A, 0, A.row_sizes.size, 0), _axis_sums(B, 0, B.col_sizes.size, 1)))
            selfcheck = {"property": "sum(C) == colsum(A) . rowsum(B)", "rel_err": abs(got - exp) / max(abs(exp), 1e-300)}
            # leave the double-buffering state as one_step expects it: buffer of the next step zeroed
            acc.memset_zero(d_cs[step_no[0] % 2], zs)
            acc.event_record(ev_zero[step_no[0] % 2], zs)
            acc.stream_sync(zs)
            acc.stream_sync(s)
        except Exception as ex:
            selfcheck = {"error": repr(ex)[:200]}

    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.3)
    launches0 = acc.launch_count()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    with torch.cuda.stream(tstream):
        for k in range(args.steps):
            ev[k][0].record(tstream)
            one_step()
            ev[k][1].record(tstream)
    acc.stream_sync(s)
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    launches = acc.launch_count() - launches0
    clocks = sampler.stop()
    # kernel-only time of one drain (no memset in flight), for the roofline of the dominant kernel
    kev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(3)]
    with torch.cuda.stream(tstream):
        for k in range(3):
            kev[k][0].record(tstream)
            drain(d_cs[0])
            kev[k][1].record(tstream)
    torch.cuda.synchronize()
    kern_ms = [kev[k][0].elapsed_time(kev[k][1]) for k in range(3)]
    step_ms = [ev[k][0].elapsed_time(ev[k][1]) for k in range(args.steps)]
    ms_per_step = float(np.mean(step_ms))
    value = flop / (ms_per_step * 1e-3) * 1e-9
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (float(np.mean(kern_ms)) * 1e-3) * 1e-9
    traffic = None
    ncu_json = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(ncu_json):
        try:
            traffic = json.load(open(ncu_json)).get(args.config, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    if bf16:
        tile = acc.bf16_tile_bytes(int(w["sizes"][0]), int(w["sizes"][0]))
        runs = sum(1 + int(np.count_nonzero(st["dev"][1:, 2] != st["dev"][:-1, 2])) for st in stacks)
        alg_bytes = n_entries * (2 * tile + 12) + runs * 2 * 4 * int(w["sizes"][0]) ** 2
        achieved = alg_bytes / (float(np.mean(kern_ms)) * 1e-3) * 1e-9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "kernel": "smm_dmma_kernel<%s> (x%d launches/step)" % (",".join(str(x) for x in (stacks[0]["m"], stacks[0]["n"], stacks[0]["k"])), len(stacks)),
                "algorithmic_bytes_per_launch": alg_bytes / max(len(stacks), 1), "avg_launch_us": float(np.mean(kern_ms)) * 1e3 / max(len(stacks), 1),
                "kernel_only_gflops": flop / (float(np.mean(kern_ms)) * 1e-3) * 1e-9, "fp64_tensor_peak_gflops_measured": 37050.0}
    if not bf16:
        # the streaming-HBM model is not what binds this kernel (operands are re-used out of L2: frac > 1); the FP64 tensor pipe
        # (DMMA.8x8x4, measured 37.05 TFLOP/s, profiles/microbench_r01.txt) and its padded ceiling (tiles of 8x8x4) are quoted beside it
        m0, n0, k0 = stacks[0]["m"], stacks[0]["n"], stacks[0]["k"]
        pad = (m0 * n0 * k0) / float(((m0 + 7) // 8 * 8) * ((n0 + 7) // 8 * 8) * ((k0 + 3) // 4 * 4))
        roofline["alt_bounds"] = {"frac_of_fp64_tensor_peak": roofline["kernel_only_gflops"] / 37050.0,
                                  "frac_of_padded_fp64_tensor_ceiling": roofline["kernel_only_gflops"] / (37050.0 * pad),
                                  "mean_run_length": n_entries / max(1, sum(1 + int(np.count_nonzero(st["dev"][1:, 2] != st["dev"][:-1, 2])) for st in stacks)),
                                  "dram_traffic_frac_of_hbm_peak": (traffic / (roofline["avg_launch_us"] * 1e-6) * 1e-9 / peak) if traffic else None,
                                  "traffic_note": "dram bytes per launch from the ncu --set full capture of the RED-flush kernel (profiles/ncu_summary.json)"}
    if bf16:
        try:
            tpeak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]) * 1e3
        except Exception:
            tpeak = 1590e3
        useful = roofline["kernel_only_gflops"]
        issued = useful * (128.0 * 32 * 32) / (23.0 ** 3)  # per entry two tcgen05.mma of M=128,N=32,K=16
        roofline.update({"kernel": "smm_bf16_kernel (tcgen05.mma M128 N32 K16, x%d launches/step)" % len(stacks),
                         "tensor": {"peak_gflops": tpeak, "useful_frac": useful / tpeak, "issued_frac": issued / tpeak,
                                    "note": "useful = 2*23^3 per product; issued = padded MMA flops (M128 x N32 x K32 per product)"}})

    # ---- end to end through the host engine, host buffers pinned
    e2e = None
    if not args.no_e2e and not bf16:
      try:
          nthreads = args.threads or min(32, max(1, (os.cpu_count() or 2) // 2))
          for d in d_cs:
              d.free()
          pa = acc.host_alloc((A.data.size,), np.float64)
          pb = acc.host_alloc((B.data.size,), np.float64)
          pa.array[:] = A.data
          pb.array[:] = B.data
          cfg_e2e = host.default_cfg(n_stacks=n_st, row_chunks=args.row_chunks)
          dm = DeviceMultiply(acc, bs, bs, bs, A.data.size, B.data.size, B.nblks, nthreads=nthreads, cfg=cfg_e2e)
          a_l, b_l = A.list3(), B.list3()
          pcs = None
          times = []
          for it in range(max(1, args.e2e_warmup) + args.e2e_steps):
              acc.device_synchronize()
              t0 = time.perf_counter()
              dm.upload_panels(pa.array, pb.array, b_l, a_list3=a_l if args.pipelined_upload else None)
              t_up = time.perf_counter()
              dm.multiply(a_l, b_l)
              t_mul = time.perf_counter()
              if pcs is None:  # pinned result buffers, sized at the first pass (pooled afterwards, like DBCSR's memory pools)
                  dm.engine.sync()
                  pcs = [acc.host_alloc((max(dm.engine.c_capacity(t), 1),), np.float64) for t in range(nthreads)]
                  prod = dm.download_c([p.array for p in pcs])
                  dm.set_result_buffers([p.array for p in pcs])
              else:
                  prod = dm.download_c()
              dt = time.perf_counter() - t0
              if it >= max(1, args.e2e_warmup):
                  times.append(dt)
                  phases = {"enqueue_upload_ms": (t_up - t0) * 1e3, "host_build_and_enqueue_ms": (t_mul - t_up) * 1e3,
                            "drain_and_d2h_ms": (time.perf_counter() - t_mul) * 1e3}
          stack_bytes = 12 * n_entries
          e2e = {"value": flop / float(np.mean(times)) * 1e-9, "unit": "GFLOP/s", "h2d_bytes_per_step": int(dm.h2d_bytes + stack_bytes),
                 "d2h_bytes_per_step": int(dm.d2h_bytes), "ms_per_step": float(np.mean(times)) * 1e3, "host_threads": nthreads, "row_chunks_per_thread": args.row_chunks, "pipelined_upload": bool(args.pipelined_upload),
                 "host_build_seconds": dm.engine.build_seconds(), "c_blocks": prod.nblks, "phases_last_step": phases, "timing": "wall clock around the public call, device synchronised on both sides"}
          dm.close()
          for p in [pa, pb] + pcs:
              p.free()

      except Exception as ex:  # the headline line must still be printed (e.g. not enough pinned memory on this host)
        e2e = {"value": None, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": repr(ex)[:300]}

    cpu = None
    if not args.no_cpu and not bf16:
        try:
            cpu = cpu_reference_sample(w, args.ref_entries, n_stacks=n_st)
        except Exception as ex:
            cpu = {"value": None, "unit": "GFLOP/s", "cores": 0, "kind": "port", "sample": "failed: " + repr(ex)[:200]}

    out = {"metric": METRIC_NAMES[args.config],
           "value": value, "unit": "GFLOP/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16" if bf16 else "f64", "data": "synthetic (numpy PCG64 seed 42, uniform(0,1))",
           "config": workload_config(w, {"products": n_entries, "flop": flop, "stacks": len(stacks), "c_blocks": int(c_nblks),
                                         "timed": "CUDA events on the launching stream; step = %d libsmm_acc_process calls into a zeroed C buffer + the memset of the next step's (pooled, double-buffered) C buffer on a side stream, joined before the step ends" % len(stacks)}),
           "clocks": clocks, "gpu_launches": int(launches), "wall_s_timed_region": t_wall, "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu,
           "selfcheck": selfcheck}
    print(json.dumps(out))
    for d in (d_a, d_b, d_st):
        d.free()
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
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-warmup", type=int, default=1)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-selfcheck", action="store_true", help="skip the full-size sum(C) check before the timed region")
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
