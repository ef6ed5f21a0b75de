"""CPU dry runs of bench.py's GPU arms with the device mocked away (fake Acc, fake torch.cuda / torch.distributed): they cannot
say anything about numbers, but they execute every Python statement of run_single / cannon.bench_main (self-checks, double
buffering, roofline / alt_bounds, JSON assembly) so that a typo cannot cost the round's only GPU run of the bench."""
import contextlib
import io
import json
import os
import types
import unittest.mock as um

import numpy as np
import torch
import torch.distributed as dist

import bench
from dbcsr_b200 import cannon, host
from dbcsr_b200 import lib as acclib


class FakeEvent:
    def __init__(self, enable_timing=False):
        pass

    def record(self, s=None):
        pass

    def elapsed_time(self, o):
        return 2.0

    def synchronize(self):
        pass


class FakeStream:
    def __init__(self, priority=0):
        pass

    def wait_event(self, e):
        pass

    def synchronize(self):
        pass


class FakeDev:
    _next = 1 << 20

    def __init__(self, nbytes):
        self.nbytes, self.ptr = nbytes, FakeDev._next
        FakeDev._next += (nbytes + 255) // 256 * 256

    def free(self):
        pass


class FakeAcc:
    def __init__(self, dev=0):
        self.n, self._keep = 0, []

    def stream_create(self, name, prio):
        buf = acclib.ctypes.create_string_buffer(8)
        self._keep.append(buf)
        return acclib.ctypes.addressof(buf)

    def to_device(self, arr, s):
        return FakeDev(np.asarray(arr).nbytes)

    def dev_alloc(self, n):
        return FakeDev(n)

    def to_host(self, dev, shape, dtype, s):
        return np.zeros(shape, dtype=dtype)

    def process(self, *a, **k):
        self.n += 1
        return 0

    def launch_count(self):
        return self.n

    def event_create(self):
        return object()

    def fp64_peak_gflops(self, s):
        return 37000.0

    def fp64_peak_sustained_gflops(self, s, seconds=0.4):
        return 33000.0

    def bf16_rk_tile_bytes(self, rows):
        return (rows + 7) // 8 * 512

    def bf16_rk_slot_bytes(self, rows, b_operand):
        return 2048 if b_operand else (rows + 7) // 8 * 512

    def __getattr__(self, name):  # stream_sync, memset_zero, event_record, stream_wait_event, stream_destroy, ...
        return lambda *a, **k: None


FAKE_CUDA = types.SimpleNamespace(Event=FakeEvent, Stream=FakeStream, ExternalStream=lambda p: FakeStream(), synchronize=lambda: None,
                                  stream=lambda s: contextlib.nullcontext(), set_device=lambda d: None, current_stream=lambda: FakeStream())


def _common_patches():
    return [um.patch.object(torch, "cuda", FAKE_CUDA), um.patch.object(acclib, "Acc", FakeAcc),
            um.patch.object(bench.ClockSampler, "start", lambda self: None),
            um.patch.object(bench.ClockSampler, "stop", lambda self: {"sm_mhz": None, "sm_max_mhz": None, "reasons": []})]


def test_run_single_dry_run():
    args = types.SimpleNamespace(config="cfg2", nblk=40, warmup=3, steps=2, no_e2e=True, no_cpu=True, no_selfcheck=False, threads=0,
                                 row_chunks=4, pipelined_upload=False, e2e_steps=1, e2e_warmup=1, ref_entries=1000, gpus=1, impl="ours",
                                 no_chain=False, probe_blocks=50, no_extra=False, extra_nblk=24, no_gpu_baseline=True, cfg3_streams=3, no_clock_sampler=False, dev_tile=0, no_tiled=True, tiled_sweep=False, no_peak_probes=False,
                                 e2e_builder="host", dev_threads=1, dev_row_chunks=2)
    out = io.StringIO()
    with contextlib.ExitStack() as st:
        for p in _common_patches() + [um.patch.object(host, "transpose_panel", lambda *a, **k: None)]:
            st.enter_context(p)
        with contextlib.redirect_stdout(out):
            bench.run_single(args)
    d = json.loads(out.getvalue().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "clocks", "gpu_launches", "roofline", "e2e", "cpu_baseline", "selfcheck"):
        assert key in d, key
    assert d["gpu_launches"] > 0 and d["selfcheck"]["rel_err"] == 1.0  # the fake device returns zeros: the check really compares
    assert d["selfcheck"]["probed_blocks"] == 50 and d["selfcheck"]["probe_max_rel_err"] == 1.0 and d["selfcheck"]["ok"] is False
    r = d["roofline"]
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic", "alt_bounds")) <= set(r) and r["bound"] == "tensor" and r["unit"] == "TFLOP/s"
    assert r["alt_bounds"]["hbm_streaming_model"]["mean_run_length"] >= 1.0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    x = d["extra_configs"]
    assert x["cfg3"]["value"] > 0 and x["cfg3"]["selfcheck"]["probed_blocks"] > 0 and "mixed" in x["cfg3"]["metric"]
    assert x["cfg4"]["value"] > 0 and x["cfg4"]["dtype"] == "bf16" and x["cfg4"]["roofline"]["useful_over_issued"] > 0
    assert x["cfg4"]["selfcheck"]["probe_max_rel_err"] == 1.0


def test_cannon_bench_main_dry_run():
    orig = cannon.CannonMultiply

    class CpuCannon(orig):
        def __init__(self, w, rank, world, device, acc=None, nthreads=1, **kw):
            orig.__init__(self, w, rank, world, "cpu", acc=None, nthreads=1, mode=host.RECORD)
            self.acc = acc

    real_tensor = torch.tensor

    def tensor_cpu(*a, **k):
        k.pop("device", None)
        return real_tensor(*a, **k)

    def fake_exit(code):
        raise SystemExit(code)

    args = types.SimpleNamespace(config="cfg2", nblk=40, warmup=3, steps=2, no_e2e=True, no_cpu=True, threads=0, e2e_steps=1, e2e_warmup=1, gpus=1,
                                 no_selfcheck=False, probe_blocks=40)
    out = io.StringIO()
    env = dict(RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", DBCSR_B200_REPLAY="py")  # the C replay creates CUDA streams; the Python loop is mocked
    with contextlib.ExitStack() as st:
        for p in _common_patches() + [um.patch.object(torch, "tensor", tensor_cpu), um.patch.object(cannon, "CannonMultiply", CpuCannon),
                                      um.patch.object(dist, "init_process_group", lambda *a, **k: None),
                                      um.patch.object(dist, "barrier", lambda *a, **k: None),
                                      um.patch.object(dist, "all_reduce", lambda *a, **k: None), um.patch.object(os, "_exit", fake_exit),
                                      um.patch.dict(os.environ, env)]:
            st.enter_context(p)
        with contextlib.redirect_stdout(out):
            try:
                cannon.bench_main(args)
            except SystemExit:
                pass
    d = json.loads([ln for ln in out.getvalue().splitlines() if ln.startswith("{")][-1])
    for key in ("metric", "value", "n_gpus", "ms_per_step", "scaling", "config", "gpu_launches", "roofline", "e2e", "selfcheck", "exchange"):
        assert key in d, key
    assert "isolated_ms_per_step" in d["config"] and "back to back" in d["config"]["timed"]  # the contract's region timing
    # zeros from the fake device: sum error 1.0, every probed block off by 1.0 (weighted x10 in the combined figure)
    assert d["selfcheck"]["rel_err"] == 10.0 and d["selfcheck"]["probe_max_rel_err"] == 1.0 and d["selfcheck"]["probed_blocks"] > 0
    assert d["selfcheck"]["ok"] is False and d["scaling"] == "strong" and "peer pull" in d["config"]["parallelism"] or "NCCL" in d["config"]["parallelism"]
