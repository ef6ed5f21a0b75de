"""CPU tests of the multi-GPU path (N>1): the Cannon schedule arithmetic, and the whole distributed multiply on the gloo backend
(world_size 2, 4 and 8 = grids 1x2, 2x2, 2x4) with the recorded host stacks drained by the oracle and compared with the
oracle's global block product."""
import os
import socket

import numpy as np
import pytest

from dbcsr_b200 import cannon, workload


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_schedule_covers_every_slice_once_and_messages_match(world):
    sc = cannon.Schedule(world)
    for r in range(world):
        assert sorted(sc.slice_at(r, t) for t in range(sc.V)) == list(range(sc.V))
    for t in range(sc.V):
        sends = {}
        for r in range(world):
            _, _, snd = sc.transfers(r, t)
            for dst, kind, s in snd:
                assert (dst, kind) not in sends  # at most one panel of a kind per destination and tick
                sends[(dst, kind)] = (r, s)
        for r in range(world):
            ra, rb, _ = sc.transfers(r, t)
            s = sc.slice_at(r, t)
            for kind, src in (("a", ra), ("b", rb)):
                if src is None:
                    assert (r, kind) not in sends
                else:
                    assert sends.pop((r, kind)) == (src, s)
        assert not sends


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nblk, sizes, q, device_build=False):
    import torch.distributed as dist

    from oracle import oracle as orc

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)
    bs = workload.block_sizes(nblk, sizes, rng)
    A = workload.random_panel(bs, bs, 0.25, rng)
    B = workload.random_panel(bs, bs, 0.25, rng)
    w = dict(name="t", nblk=nblk, sizes=sizes, occupation=0.25, m_sizes=bs, n_sizes=bs, k_sizes=bs, A=A, B=B, seed=123)
    from dbcsr_b200 import host

    # device_build: the stacks of every tick come from the device-side builder's passes (run in host loops without a GPU); the k block
    # sizes change from tick to tick (set_k_sizes) and every tick accumulates onto the C index of the earlier ones
    cm = cannon.CannonMultiply(w, rank, world, "cpu", acc=None, nthreads=1, cfg=host.default_cfg(mm_stack_size=200, n_stacks=max(3, len(sizes))),
                               mode=host.RECORD | (host.DEVICE_BUILD if device_build else 0))
    # drain the recorded host stacks of every tick with the oracle's CPU path on the panels this rank held at that tick
    sc = cm.sched
    per_tick = []
    V = sc.V
    pending = cm.post_exchange(0)
    rows, cols, blk_p, ds = None, None, None, 0
    c = None
    n_before = 0
    for t in range(V):
        for wk in pending:
            wk.wait()
        pending = cm.post_exchange(t + 1) if t + 1 < V else []
        (abuf, anb, anz), (bbuf, bnb, bnz) = cm.panel_of_tick(t, "a"), cm.panel_of_tick(t, "b")
        a_idx, b_idx = cm.index_to_host(abuf, anb, anz), cm.index_to_host(bbuf, bnb, bnz)
        cm.engine.set_k_sizes(cm.k_sizes[sc.slice_at(rank, t)])
        cm.engine.multiply(a_idx, None, b_idx, None)
        st = cm.engine.stacks()
        a_data = abuf[:anz * 8].numpy().view(np.float64).copy()
        b_data = bbuf[:bnz * 8].numpy().view(np.float64).copy()
        per_tick.append((st[n_before:], a_data, b_data))
        n_before = len(st)
    rows, cols, blk_p, ds = cm.engine.c_index(0)
    assert cm.engine.device_built_ticks == (V if device_build else 0)
    c = np.zeros(max(ds, 1))
    for stacks, a_data, b_data in per_tick:
        for s_ in stacks:
            orc.host_stack(s_["host"], a_data if a_data.size else np.zeros(1), b_data if b_data.size else np.zeros(1), c)
    # global coordinates of this rank's C blocks
    out = {}
    for r, cc, o in zip(rows, cols, blk_p):
        gr, gc = int(r) + cm.rsp[cm.i], int(cc) + cm.csp[cm.j]
        m, n = int(bs[gr - 1]), int(bs[gc - 1])
        out[(gr, gc)] = c[o - 1:o - 1 + m * n].copy()
    q.put((rank, out, cm.engine.flop()))
    dist.barrier()
    cm.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,sizes,device_build", [(2, [23], False), (4, [5, 13, 23], False), (8, [23], False), (4, [5, 13, 23], True), (8, [23], True)],
                         ids=["2", "4_mixed", "8", "4_mixed_device_builder", "8_device_builder"])
def test_distributed_multiply_gloo(world, sizes, device_build):
    import torch.multiprocessing as mp

    from oracle import oracle as orc

    nblk = 24
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nblk, sizes, q, device_build)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(123)
    bs = workload.block_sizes(nblk, sizes, rng)
    A = workload.random_panel(bs, bs, 0.25, rng)
    B = workload.random_panel(bs, bs, 0.25, rng)
    Cref = orc.multiply_blocks(orc.BlockMatrix(bs, bs, A.rows, A.cols, data=A.data), orc.BlockMatrix(bs, bs, B.rows, B.cols, data=B.data))
    got = {}
    flop = 0
    for rank, out, f in results:
        assert not (set(out) & set(got))  # every C block is owned by exactly one rank
        got.update(out)
        flop += f
    assert set(got) == set(zip(Cref.rows.tolist(), Cref.cols.tolist()))
    num = den = 0.0
    for r, c, o in zip(Cref.rows, Cref.cols, Cref.offsets):
        m, n = int(bs[r - 1]), int(bs[c - 1])
        ref = Cref.data[o:o + m * n]
        num += float(((got[(int(r), int(c))] - ref) ** 2).sum())
        den += float((ref ** 2).sum())
    assert (num / den) ** 0.5 <= 1e-12
    assert flop == sum(2 * int(bs[r - 1]) * int(bs[c - 1]) * int(bs[B.cols[B.rows == c] - 1].sum()) for r, c in zip(A.rows, A.cols))


def _images_worker(rank, world, port, nblk, sizes, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(321)
    bs = workload.block_sizes(nblk, sizes, rng)
    A = workload.random_panel(bs, bs, 0.3, rng)
    B = workload.random_panel(bs, bs, 0.3, rng)
    sc = cannon.Schedule(world)
    # a 2-d block distribution unrelated to the Cannon layout (DBCSR: random / round-robin row_dist, col_dist)
    row_dist = rng.integers(0, sc.pr, nblk)
    col_dist = rng.integers(0, sc.pc, nblk)
    w_rep = dict(name="t", nblk=nblk, sizes=sizes, occupation=0.3, m_sizes=bs, n_sizes=bs, k_sizes=bs, A=A, B=B, seed=321)
    w_dist = dict(w_rep)
    w_dist["A_dist"] = cannon.DistMatrix.from_global(A, row_dist, col_dist, sc, rank)
    w_dist["B_dist"] = cannon.DistMatrix.from_global(B, row_dist, col_dist, sc, rank)
    from dbcsr_b200 import host

    cfg = host.default_cfg(mm_stack_size=200, n_stacks=max(3, len(sizes)))
    cm_rep = cannon.CannonMultiply(w_rep, rank, world, "cpu", acc=None, nthreads=1, cfg=cfg)
    cm_dist = cannon.CannonMultiply(w_dist, rank, world, "cpu", acc=None, nthreads=1, cfg=cfg)
    same = sorted(cm_rep.home) == sorted(cm_dist.home)
    for key in cm_rep.home:
        a, b = cm_rep.home[key], cm_dist.home.get(key)
        same = same and b is not None and np.array_equal(a.rows, b.rows) and np.array_equal(a.cols, b.cols) and np.array_equal(a.data, b.data) \
            and np.array_equal(a.row_sizes, b.row_sizes) and np.array_equal(a.col_sizes, b.col_sizes)
    owned = (w_dist["A_dist"].panel.nblks, w_dist["B_dist"].panel.nblks)
    st_rep, st_dist = cm_rep.run(), cm_dist.run()
    same_stacks = len(st_rep) == len(st_dist) and all(
        len(x) == len(y) and all(np.array_equal(p["host"], r["host"]) for p, r in zip(x, y)) for x, y in zip(st_rep, st_dist))
    q.put((rank, bool(same), bool(same_stacks), owned, cm_dist.engine.flop()))
    dist.barrier()
    cm_rep.close()
    cm_dist.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,sizes", [(2, [23]), (8, [5, 13, 23])])
def test_make_images_from_distributed_input_gloo(world, sizes):
    """make_images: ranks that only hold their own blocks of A and B (2-d block distribution) exchange them in one all-to-all and
    end up with exactly the home panels (Cannon's initial alignment, panel-local BCSR order) that slicing a replicated matrix
    gives, hence identical stacks in the multiply."""
    import torch.multiprocessing as mp

    nblk = 24
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_images_worker, args=(r, world, port, nblk, sizes, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(321)
    bs = workload.block_sizes(nblk, sizes, rng)
    A = workload.random_panel(bs, bs, 0.3, rng)
    B = workload.random_panel(bs, bs, 0.3, rng)
    assert sum(r[3][0] for r in results) == A.nblks and sum(r[3][1] for r in results) == B.nblks  # every block owned once
    for rank, same, same_stacks, owned, flop in results:
        assert same and same_stacks, rank
    assert sum(r[4] for r in results) > 0


def test_replay_and_pipelined_run_control_flow_with_mocked_cuda():
    """The GPU-only driver paths of the Cannon bench (replay_step with one receive buffer per tick and all pulls posted up front;
    _run_prefetch_all = pipelined multi-tick engine) dry-run on CPU for all 8 ranks of the 2x4 grid: CUDA streams / events are
    mocked, peer access is emulated by handing every rank the other ranks' home buffers, libsmm_acc_process is a counter.
    Checks: no Python error, every tick launches its recorded stacks, and after a step the buffer of every (tick, kind) holds
    exactly the home panel of the slice that tick multiplies."""
    import contextlib
    import types
    import unittest.mock as um

    import torch

    from dbcsr_b200 import host

    class FakeEvent:
        def __init__(self, enable_timing=False):
            pass

        def record(self, stream=None):
            pass

        def synchronize(self):
            pass

    class FakeStream:
        def __init__(self, priority=0):
            pass

        def wait_event(self, e):
            assert isinstance(e, FakeEvent)

        def synchronize(self):
            pass

    fake_cuda = types.SimpleNamespace(Event=FakeEvent, Stream=FakeStream, stream=lambda s: contextlib.nullcontext(),
                                      current_stream=lambda: FakeStream(), synchronize=lambda: None)

    class FakeTorch:
        def __getattr__(self, name):
            return fake_cuda if name == "cuda" else getattr(torch, name)

    launched = []

    class FakeAcc:
        def process(self, host7, dev, S, a, b, c, m, n, k, dm, s1, s2):
            launched.append(S)
            return 0

    world = 8
    sc = cannon.Schedule(world)
    w = workload.make_config("cfg2", nblk=48)
    cms = []
    with um.patch("torch.distributed.all_reduce", lambda t, op=None: None):
        for r in range(world):
            cms.append(cannon.CannonMultiply(w, r, world, "cpu", acc=None, nthreads=1))
    meta = sum(torch.from_numpy(cm.meta) for cm in cms).numpy()  # what the all_reduce of the panel sizes would have produced
    max_bytes = int((meta[..., 1] * 8 + meta[..., 0] * 12).max())
    orig_empty = torch.empty
    for r, cm in enumerate(cms):
        cm.meta = meta
        cm.peer_buf = {(q,) + key: t for q, other in enumerate(cms) if q != r for key, t in other.home_buf.items()}
        cm.prefetch_all, cm.nbuf = True, sc.V
        cm.recv = {kind: [torch.empty(max(max_bytes, 16), dtype=torch.uint8) for _ in range(cm.nbuf)] for kind in "ab"}
        cm.torch = FakeTorch()
        assert cm.tick_order()[0] == 0  # Cannon's initial alignment: the first tick needs no pull on any rank
        ra, rb, _ = sc.transfers(r, 0)
        assert ra is None and rb is None
        per_tick = cm.run()  # record mode: tick-by-tick loop with (emulated) peer copies
        assert len(per_tick) == sc.V
        cm.replay = [[(0, st["dev"].shape[0], st["max_m"], st["max_n"], st["max_k"], st["defined_mnk"]) for st in tick] for tick in per_tick]
        cm.replay_stacks = torch.zeros(8, dtype=torch.int32)
        cm.replay_cs = [torch.zeros(4), torch.zeros(4)]
        cm.zero_stream, cm.ev_zero, cm.ev_free = FakeStream(), [FakeEvent(), FakeEvent()], [FakeEvent(), FakeEvent()]
        cm.step_no, cm.cs, cm.cs_torch, cm.comm_stream = 0, 0, FakeStream(), FakeStream()
        cm.acc = FakeAcc()
        for kind in "ab":  # poison the receive buffers: the replay has to refill them
            for b in cm.recv[kind]:
                b.fill_(255)
        n0 = len(launched)
        for _ in range(3):
            cm.replay_step()
        assert len(launched) - n0 == 3 * sum(len(x) for x in cm.replay)
        for t in range(sc.V):
            for kind in "ab":
                buf, nblk, nze = cm.panel_of_tick(t, kind)
                s = sc.slice_at(r, t)
                src = cm.panel_meta(kind, s, cm.i, cm.j)[0]
                n = nze * 8 + nblk * 12
                assert torch.equal(buf[:n], cms[src].home_buf[(kind, s)][:n]), (r, t, kind)
    # pipelined engine path on two ranks (the engine records instead of launching)
    for cm in cms[:2]:
        cm.engine.reset()
        cm.mode = host.RECORD
        with um.patch("torch.empty", lambda *a, **k: orig_empty(*a, **{x: y for x, y in k.items() if x != "pin_memory"})):
            per_tick = cm._run_prefetch_all()
        assert len(per_tick) == sc.V and cm.flop > 0
    for cm in cms:
        cm.acc = None
        cm.close()
