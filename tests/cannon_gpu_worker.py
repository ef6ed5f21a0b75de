"""Worker of tests/test_gpu_cannon.py: one rank of a multi-process Cannon multiply on real GPUs (launched by torch.distributed.run).
Runs dbcsr_b200.cannon.CannonMultiply in LAUNCH mode (host engine builds and uploads stacks every tick) and through the replay
path (pre-built device stacks, what bench.py --gpus N times), downloads this rank's C and compares it BLOCK BY BLOCK with the
oracle's product of the global matrices (orc.multiply_blocks), like the reference checks every distributed multiply against a
dense DGEMM (tests/dbcsr_test_multiply.F:629-630,753-759, run with mpiexec -np 2).
With fewer GPUs than ranks the ranks share device 0 (gloo for the set-up collectives, CUDA IPC peer pull for the panels)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def check_blocks(tag, got, ref, row0, row1, col0, col1, tol=1e-10):
    """got: {(global_row, global_col): flat column-major block}; ref: same for the whole product."""
    mine = {k: v for k, v in ref.items() if row0 < k[0] <= row1 and col0 < k[1] <= col1}
    assert set(got) == set(mine), "%s: block pattern differs (%d vs %d blocks)" % (tag, len(got), len(mine))
    worst = 0.0
    for k, v in mine.items():
        den = float(np.linalg.norm(v))
        err = float(np.linalg.norm(got[k] - v)) / max(den, 1e-300)
        worst = max(worst, err)
    assert worst <= tol, "%s: worst block error %.3e" % (tag, worst)
    return len(mine), worst


def main():
    import torch
    import torch.distributed as dist

    from dbcsr_b200 import cannon, host, workload
    from dbcsr_b200 import lib as acclib
    from oracle import oracle as orc

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    ndev = torch.cuda.device_count()
    dev = rank % ndev
    torch.cuda.set_device(dev)
    shared_gpu = ndev < world
    if shared_gpu:
        dist.init_process_group("gloo")
        os.environ["DBCSR_B200_EXCHANGE"] = "p2p"
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    acc = acclib.Acc(dev)
    report = {"rank": rank, "world": world, "shared_gpu": shared_gpu, "cases": []}
    for cfg_name, nblk in (("cfg2", int(os.environ.get("CANNON_TEST_NBLK", "64"))), ("cfg3", 48)):
        w = workload.make_config(cfg_name, nblk=nblk)
        A, B, bs = w["A"], w["B"], w["m_sizes"]
        Cref = orc.multiply_blocks(orc.BlockMatrix(A.row_sizes, A.col_sizes, A.rows, A.cols, data=A.data),
                                   orc.BlockMatrix(B.row_sizes, B.col_sizes, B.rows, B.cols, data=B.data))
        ref = {}
        for r, c, o in zip(Cref.rows, Cref.cols, Cref.offsets):
            ref[(int(r), int(c))] = Cref.data[o:o + int(bs[r - 1]) * int(bs[c - 1])]
        # ---- engine path (LAUNCH): stacks built by the host threads and uploaded every tick
        cm = cannon.CannonMultiply(w, rank, world, "cuda:%d" % dev, acc=acc, nthreads=2, cfg=host.default_cfg(mm_stack_size=1000, n_stacks=3 if len(w["sizes"]) <= 3 else len(w["sizes"])))
        r0, r1, c0, c1 = cm.rsp[cm.i], cm.rsp[cm.i + 1], cm.csp[cm.j], cm.csp[cm.j + 1]
        for repeat in range(2):  # the second multiply runs on pooled resources (reset), like consecutive multiplies in DBCSR
            if repeat:
                cm.engine.reset()
            cm.run()
            cm.engine.sync()
            torch.cuda.synchronize()
            got = {}
            for t in range(cm.engine.nthreads):
                rows, cols, blk_p, ds = cm.engine.c_index(t)
                if ds == 0:
                    continue
                buf = np.empty(ds)
                cm.engine.c_to_host(t, buf)
                for rr, cc, p in zip(rows, cols, blk_p):
                    nz = int(cm.m_sizes[rr - 1]) * int(cm.n_sizes[cc - 1])
                    key = (int(rr) + r0, int(cc) + c0)
                    assert key not in got, "block %r produced by two threads" % (key,)
                    got[key] = buf[p - 1:p - 1 + nz].copy()
            nb, worst = check_blocks("%s engine rank %d pass %d" % (cfg_name, rank, repeat), got, ref, r0, r1, c0, c1)
        report["cases"].append({"config": cfg_name, "path": "engine (exchange: %s)" % ("peer pull" if cm.peer_buf is not None else "nccl send/recv"),
                                "prefetch_all": bool(cm.prefetch_all), "blocks": nb, "worst_rel_err": worst})
        # ---- replay path: pre-built device stacks, what bench.py --gpus N times
        cm.build_replay()
        for repeat in range(3):
            cm.replay_step()
        torch.cuda.synchronize()
        c = cm.replay_c.cpu().numpy()
        rows, cols, blk_p = cm.replay_c_index
        got = {}
        for rr, cc, p in zip(rows, cols, blk_p):
            nz = int(cm.m_sizes[rr - 1]) * int(cm.n_sizes[cc - 1])
            got[(int(rr) + r0, int(cc) + c0)] = c[p - 1:p - 1 + nz]
        nb, worst = check_blocks("%s replay rank %d" % (cfg_name, rank), got, ref, r0, r1, c0, c1)
        report["cases"].append({"config": cfg_name, "path": "replay", "blocks": nb, "worst_rel_err": worst})
        dist.barrier()
        cm.close()
    # ---- distributed input (row a3): every rank holds only ITS blocks under a 2-d block distribution unrelated to the Cannon layout;
    #      make_images moves them to their home panels (all-to-all over NCCL, device tensors), then the same multiply
    w = workload.make_config("cfg2", nblk=int(os.environ.get("CANNON_TEST_NBLK", "64")))
    A, B, bs = w["A"], w["B"], w["m_sizes"]
    Cref = orc.multiply_blocks(orc.BlockMatrix(A.row_sizes, A.col_sizes, A.rows, A.cols, data=A.data),
                               orc.BlockMatrix(B.row_sizes, B.col_sizes, B.rows, B.cols, data=B.data))
    ref = {(int(r), int(c)): Cref.data[o:o + int(bs[r - 1]) * int(bs[c - 1])] for r, c, o in zip(Cref.rows, Cref.cols, Cref.offsets)}
    sc = cannon.Schedule(world)
    rng = np.random.default_rng(1234)  # the same maps on every rank
    row_dist, col_dist = rng.integers(0, sc.pr, w["nblk"]), rng.integers(0, sc.pc, w["nblk"])
    w_dist = {k: v for k, v in w.items() if k not in ("A", "B")}
    w_dist["A_dist"] = cannon.DistMatrix.from_global(A, row_dist, col_dist, sc, rank)
    w_dist["B_dist"] = cannon.DistMatrix.from_global(B, row_dist, col_dist, sc, rank)
    cm = cannon.CannonMultiply(w_dist, rank, world, "cuda:%d" % dev, acc=acc, nthreads=2, cfg=host.default_cfg(mm_stack_size=1000))
    r0, r1, c0, c1 = cm.rsp[cm.i], cm.rsp[cm.i + 1], cm.csp[cm.j], cm.csp[cm.j + 1]
    cm.run()
    cm.engine.sync()
    torch.cuda.synchronize()
    got = {}
    for t in range(cm.engine.nthreads):
        rows, cols, blk_p, ds = cm.engine.c_index(t)
        if ds == 0:
            continue
        buf = np.empty(ds)
        cm.engine.c_to_host(t, buf)
        for rr, cc, p in zip(rows, cols, blk_p):
            nz = int(cm.m_sizes[rr - 1]) * int(cm.n_sizes[cc - 1])
            got[(int(rr) + r0, int(cc) + c0)] = buf[p - 1:p - 1 + nz].copy()
    nb, worst = check_blocks("distributed input rank %d" % rank, got, ref, r0, r1, c0, c1)
    report["cases"].append({"config": "cfg2", "path": "engine, distributed input through make_images", "blocks": nb, "worst_rel_err": worst,
                            "owned_blocks": [w_dist["A_dist"].panel.nblks, w_dist["B_dist"].panel.nblks]})
    dist.barrier()
    cm.close()
    outdir = os.environ.get("CANNON_TEST_OUT")
    if outdir:
        with open(os.path.join(outdir, "rank%d.json" % rank), "w") as f:
            json.dump(report, f)
    print("CANNON_WORKER_OK " + json.dumps(report), flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
