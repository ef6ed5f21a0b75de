"""
dbcsr_b200/host.py -- ctypes binding of include/dbcsr_b200_host.h: the C++ stand-in for DBCSR's Fortran local-multiply layer
(multrec / csr stack builder / sched / accdrv).  Used by bench.py and the tests; never imports anything from oracle/.
"""
import ctypes

import numpy as np

from . import lib as acclib

_vp, _i = ctypes.c_void_p, ctypes.c_int
LAUNCH, RECORD, DEVICE_BUILD = 1, 2, 4


class Cfg(ctypes.Structure):
    _fields_ = [(n, _i) for n in ("mm_stack_size", "n_stacks", "multrec_limit", "stack_sort", "min_flop_sort", "binning_nbins",
                                  "binning_binsize", "thread_buffers", "row_chunks", "dev_tile")]


_bound = False


def _L():
    global _bound
    L = acclib.load()
    if not _bound:
        ip = ctypes.POINTER(_i)
        L.dbcsr_b200_cfg_default.argtypes = [ctypes.POINTER(Cfg)]
        L.dbcsr_b200_rec_sort_index.argtypes = [_i, _i, _i, _vp]
        L.dbcsr_b200_rec_sort_index_mt.argtypes = [_i, _i, _i, _vp, _i]
        L.dbcsr_b200_stack_sort.argtypes = [_vp, _vp, _i]
        L.dbcsr_b200_stack_binning.argtypes = [_vp, _vp, _i, _i, _i]
        L.dbcsr_b200_engine_create.argtypes = [ctypes.POINTER(Cfg), _vp, _i, _vp, _i, _vp, _i, _i, _i, ctypes.c_size_t]
        L.dbcsr_b200_engine_create.restype = _vp
        L.dbcsr_b200_engine_destroy.argtypes = [_vp]
        L.dbcsr_b200_engine_multiply.argtypes = [_vp, _vp, _i, _vp, _vp, _i, _vp]
        L.dbcsr_b200_engine_multiply_filtered.argtypes = [_vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp]
        L.dbcsr_b200_engine_set_filter.argtypes = [_vp, _vp]
        L.dbcsr_b200_row_max_epss.argtypes = [ctypes.c_double, _vp, _i, _vp]
        L.dbcsr_b200_row_max_epss.restype = None
        L.dbcsr_b200_engine_nchunks.argtypes = [_vp]
        L.dbcsr_b200_engine_chunk_rows.argtypes = [_vp, _i, ip, ip]
        L.dbcsr_b200_engine_set_chunk_events.argtypes = [_vp, _vp, _i]
        L.dbcsr_b200_engine_preset_c.argtypes = [_vp, _vp, _vp, _i, _vp, _i]
        L.dbcsr_b200_engine_set_c_symmetry.argtypes = [_vp, _i, _vp, _vp]
        L.dbcsr_b200_engine_stats.argtypes = [_vp, _vp, _i, _vp]
        L.dbcsr_b200_engine_set_host_driver.argtypes = [_vp, _vp, _vp]
        L.dbcsr_b200_engine_stats_cpu.argtypes = [_vp, _vp]
        L.dbcsr_b200_filter_index.argtypes = [ctypes.c_double, _vp, _i, _vp, _vp, _vp, _vp, ctypes.POINTER(ctypes.c_longlong)]
        L.dbcsr_b200_engine_filter_c.argtypes = [_vp, ctypes.c_double]
        L.dbcsr_b200_engine_finalize_c.argtypes = [_vp, ctypes.c_double]
        L.dbcsr_b200_finalize_index.argtypes = [_i, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(ctypes.c_longlong)]
        L.dbcsr_b200_engine_sync.argtypes = [_vp]
        L.dbcsr_b200_engine_reset.argtypes = [_vp]
        L.dbcsr_b200_engine_set_k_sizes.argtypes = [_vp, _vp, _i]
        for f in ("nthreads", "nstacks"):
            getattr(L, "dbcsr_b200_engine_" + f).argtypes = [_vp]
        for f in ("c_nblks", "c_datasize"):
            getattr(L, "dbcsr_b200_engine_" + f).argtypes = [_vp, _i]
        for f in ("c_rows", "c_cols", "c_blk_p"):
            fn = getattr(L, "dbcsr_b200_engine_" + f)
            fn.argtypes = [_vp, _i]
            fn.restype = ip
        L.dbcsr_b200_engine_c_dev.argtypes = [_vp, _i]
        L.dbcsr_b200_engine_c_dev.restype = _vp
        L.dbcsr_b200_engine_c_to_host.argtypes = [_vp, _i, _vp]
        L.dbcsr_b200_engine_c_to_host_async.argtypes = [_vp, _i, _vp]
        L.dbcsr_b200_engine_wait_event.argtypes = [_vp, _vp]
        L.dbcsr_b200_engine_set_c_host.argtypes = [_vp, _i, _vp]
        L.dbcsr_b200_engine_c_capacity.argtypes = [_vp, _i]
        L.dbcsr_b200_engine_c_capacity.restype = ctypes.c_size_t
        L.dbcsr_b200_engine_flop.argtypes = [_vp]
        L.dbcsr_b200_engine_flop.restype = ctypes.c_longlong
        L.dbcsr_b200_engine_build_seconds.argtypes = [_vp]
        L.dbcsr_b200_engine_build_seconds.restype = ctypes.c_double
        L.dbcsr_b200_engine_stack_info.argtypes = [_vp, _i, _vp]
        for f in ("stack_host", "stack_dev"):
            fn = getattr(L, "dbcsr_b200_engine_" + f)
            fn.argtypes = [_vp, _i]
            fn.restype = ip
        L.dbcsr_b200_transpose_panel.argtypes = [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]
        L.dbcsr_b200_transpose_panel_norms.argtypes = [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
        L.dbcsr_b200_replay_create.argtypes = [_i]
        L.dbcsr_b200_replay_create.restype = _vp
        L.dbcsr_b200_replay_destroy.argtypes = [_vp]
        L.dbcsr_b200_replay_destroy.restype = None
        L.dbcsr_b200_replay_set_panels.argtypes = [_vp, _i, _vp, _vp]
        L.dbcsr_b200_replay_add_pull.argtypes = [_vp, _i, _vp, _vp, ctypes.c_size_t]
        L.dbcsr_b200_replay_add_stack.argtypes = [_vp, _i, _vp, _i, _i, _i, _i, _i]
        L.dbcsr_b200_replay_set_c.argtypes = [_vp, _vp, _vp, ctypes.c_size_t, _i]
        L.dbcsr_b200_replay_current_c.argtypes = [_vp]
        L.dbcsr_b200_replay_current_c.restype = _vp
        L.dbcsr_b200_replay_step.argtypes = [_vp, _vp]
        _bound = True
    return L


def default_cfg(**kw):
    c = Cfg()
    _L().dbcsr_b200_cfg_default(ctypes.byref(c))
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def rec_sort_index(nrows, ncols, list3, depth=0):
    a = np.ascontiguousarray(list3, dtype=np.int32).reshape(-1, 3).copy()
    if depth > 0:
        _L().dbcsr_b200_rec_sort_index_mt(nrows, ncols, a.shape[0], a.ctypes.data, depth)
    else:
        _L().dbcsr_b200_rec_sort_index(nrows, ncols, a.shape[0], a.ctypes.data)
    return a


def stack_sort(params7):
    p = np.ascontiguousarray(params7, dtype=np.int32).reshape(-1, 7)
    out = np.empty((p.shape[0], 3), dtype=np.int32)
    _L().dbcsr_b200_stack_sort(p.ctypes.data, out.ctypes.data, p.shape[0])
    return out


def stack_binning(params7, nbins=4096, binsize=16):
    p = np.ascontiguousarray(params7, dtype=np.int32).reshape(-1, 7)
    out = np.empty((p.shape[0], 3), dtype=np.int32)
    _L().dbcsr_b200_stack_binning(p.ctypes.data, out.ctypes.data, p.shape[0], nbins, binsize)
    return out


def row_max_epss(filter_eps, total_row_counts):
    """row_max_epss of dbcsr_multiply's on-the-fly filter (src/mm/dbcsr_mm_cannon.F:1098-1107), float32 per block row."""
    cnt = np.ascontiguousarray(total_row_counts, dtype=np.int32)
    out = np.empty(cnt.size, dtype=np.float32)
    _L().dbcsr_b200_row_max_epss(float(filter_eps), cnt.ctypes.data, cnt.size, out.ctypes.data)
    return out


def filter_index(filter_eps, norms2, rows, cols, blk_p, nelems):
    """multrec_filtering, index part (see include/dbcsr_b200_host.h).  Returns (rows, cols, blk_p, nze_after) of the kept blocks."""
    n2 = np.ascontiguousarray(norms2, dtype=np.float64)
    r, c, p = (np.ascontiguousarray(x, dtype=np.int32).copy() for x in (rows, cols, blk_p))
    ne = np.ascontiguousarray(nelems, dtype=np.int32)
    nze = ctypes.c_longlong(0)
    kept = _L().dbcsr_b200_filter_index(float(filter_eps), n2.ctypes.data, r.size, r.ctypes.data, c.ctypes.data, p.ctypes.data, ne.ctypes.data,
                                        ctypes.byref(nze))
    return r[:kept], c[:kept], p[:kept], int(nze.value)


class Engine:
    """multrec + csr + sched + accdrv of `nthreads` host threads for one rank (see include/dbcsr_b200_host.h)."""

    def __init__(self, m_sizes, n_sizes, k_sizes, nthreads=1, mode=RECORD, cfg=None, c_capacity=0):
        self.L = _L()
        self.cfg = cfg if cfg is not None else default_cfg()
        self.m = np.ascontiguousarray(m_sizes, dtype=np.int32)
        self.n = np.ascontiguousarray(n_sizes, dtype=np.int32)
        self.k = np.ascontiguousarray(k_sizes, dtype=np.int32)
        self.h = self.L.dbcsr_b200_engine_create(ctypes.byref(self.cfg), self.m.ctypes.data, self.m.size, self.n.ctypes.data, self.n.size,
                                                 self.k.ctypes.data, self.k.size, nthreads, mode, c_capacity)
        if not self.h:
            raise acclib.AccError("dbcsr_b200_engine_create failed")

    def multiply(self, a_list3, a_dev_ptr, b_list3, b_dev_ptr, a_norms=None, b_norms=None):
        a = np.ascontiguousarray(a_list3, dtype=np.int32).reshape(-1, 3)
        b = np.ascontiguousarray(b_list3, dtype=np.int32).reshape(-1, 3)
        if a_norms is None and b_norms is None:
            rc = self.L.dbcsr_b200_engine_multiply(self.h, a.ctypes.data, a.shape[0], a_dev_ptr, b.ctypes.data, b.shape[0], b_dev_ptr)
        else:
            an = np.ascontiguousarray(a_norms, dtype=np.float32)
            bn = np.ascontiguousarray(b_norms, dtype=np.float32)
            if an.size != a.shape[0] or bn.size != b.shape[0]:
                raise ValueError("one norm per block of each panel")
            rc = self.L.dbcsr_b200_engine_multiply_filtered(self.h, a.ctypes.data, a.shape[0], a_dev_ptr, an.ctypes.data,
                                                            b.ctypes.data, b.shape[0], b_dev_ptr, bn.ctypes.data)
        if rc != 0:
            raise acclib.AccError("dbcsr_b200_engine_multiply returned %d" % rc)

    def set_filter(self, row_eps):
        """row_eps: float32 per local block row (see row_max_epss) or None to switch the on-the-fly filter off."""
        if row_eps is None:
            rc = self.L.dbcsr_b200_engine_set_filter(self.h, None)
        else:
            r = np.ascontiguousarray(row_eps, dtype=np.float32)
            if r.size != self.m.size:
                raise ValueError("one threshold per block row")
            rc = self.L.dbcsr_b200_engine_set_filter(self.h, r.ctypes.data)
        if rc != 0:
            raise acclib.AccError("dbcsr_b200_engine_set_filter failed")

    @property
    def nchunks(self):
        return self.L.dbcsr_b200_engine_nchunks(self.h)

    def chunk_rows(self, chunk):
        """Block rows (row_lo, row_hi] (1-based) of row chunk `chunk`; thread chunk % nthreads owns it."""
        lo, hi = _i(0), _i(0)
        if self.L.dbcsr_b200_engine_chunk_rows(self.h, chunk, ctypes.byref(lo), ctypes.byref(hi)) != 0:
            raise acclib.AccError("dbcsr_b200_engine_chunk_rows failed")
        return lo.value, hi.value

    def set_chunk_events(self, events):
        """events[c]: acc event recorded behind the upload of chunk c's A rows (None = no wait); used by the next multiply."""
        arr = (_vp * len(events))(*[e if e else None for e in events])
        if self.L.dbcsr_b200_engine_set_chunk_events(self.h, arr, len(events)) != 0:
            raise acclib.AccError("dbcsr_b200_engine_set_chunk_events failed")

    def preset_c(self, rows, cols, data=None, keep_sparsity=False):
        """Existing C blocks (beta*C_old, or zeros when data is None) the product accumulates onto; keep_sparsity = retain_sparsity."""
        r = np.ascontiguousarray(rows, dtype=np.int32)
        c = np.ascontiguousarray(cols, dtype=np.int32)
        d = None if data is None else np.ascontiguousarray(data, dtype=np.float64)
        if d is not None and d.size != int((self.m[r - 1].astype(np.int64) * self.n[c - 1]).sum()):
            raise ValueError("data must hold exactly the listed blocks")
        rc = self.L.dbcsr_b200_engine_preset_c(self.h, r.ctypes.data, c.ctypes.data, r.size, None if d is None else d.ctypes.data,
                                               1 if keep_sparsity else 0)
        if rc != 0:
            raise acclib.AccError("dbcsr_b200_engine_preset_c returned %d" % rc)

    def set_c_symmetry(self, on, global_rows=None, global_cols=None):
        """Product with symmetry: compute only the blocks the checkerboard rule does not store transposed
        (src/mm/dbcsr_mm_csr.F:280-292); global_rows/global_cols map local C rows/cols to global block indices (None = identity).
        Ends with the next reset()."""
        gr = None if global_rows is None else np.ascontiguousarray(global_rows, dtype=np.int32)
        gc = None if global_cols is None else np.ascontiguousarray(global_cols, dtype=np.int32)
        if (gr is not None and gr.size != self.m.size) or (gc is not None and gc.size != self.n.size):
            raise ValueError("one global index per local block row / col")
        rc = self.L.dbcsr_b200_engine_set_c_symmetry(self.h, 1 if on else 0, None if gr is None else gr.ctypes.data,
                                                     None if gc is None else gc.ctypes.data)
        if rc != 0:
            raise acclib.AccError("dbcsr_b200_engine_set_c_symmetry returned %d" % rc)

    def filter_c(self, filter_eps):
        rc = self.L.dbcsr_b200_engine_filter_c(self.h, float(filter_eps))
        if rc != 0:
            raise acclib.AccError("dbcsr_b200_engine_filter_c returned %d" % rc)

    def stats(self):
        """Scheduler statistics (DBCSR's STATISTICS table): list of dicts per (m,n,k), ordered by flop, and the totals."""
        n = self.L.dbcsr_b200_engine_stats(self.h, None, 0, None)
        table = np.zeros((max(n, 1), 7), dtype=np.int64)
        totals = np.zeros(3, dtype=np.int64)
        n = self.L.dbcsr_b200_engine_stats(self.h, table.ctypes.data, table.shape[0], totals.ctypes.data)
        rows = [dict(m=int(r[0]), n=int(r[1]), k=int(r[2]), entries=int(r[3]), stacks=int(r[4]), stacks_untuned=int(r[5]), flop=int(r[6]))
                for r in table[:n]]
        return rows, dict(flop=int(totals[0]), entries=int(totals[1]), stacks=int(totals[2]))

    def set_host_driver(self, fn):
        """Install the scheduler's host-driver route (include/dbcsr_b200_host.h): fn(thread, m, n, k, defined_mnk, params7, c_datasize)
        with params7 an (S, 7) int32 array; it must apply the stack to the caller's host work area of `thread` and return 0.
        None removes it (a stack the accelerator refuses then fails the multiply).  The library itself never computes on the CPU."""
        if fn is None:
            self._host_driver_cb = None
            self.L.dbcsr_b200_engine_set_host_driver(self.h, None, None)
            return
        proto = ctypes.CFUNCTYPE(_i, _vp, _i, _i, _i, _i, _i, ctypes.POINTER(_i), _i, _i)

        def trampoline(ctx, thread, m, n, k, defined, params, size, datasize):
            try:
                p7 = np.ctypeslib.as_array(params, shape=(size, 7)).copy()
                return int(fn(thread, m, n, k, defined, p7, datasize) or 0)
            except Exception:  # an exception must not unwind through C
                import traceback

                traceback.print_exc()
                return -51

        self._host_driver_cb = proto(trampoline)  # keep alive
        self.L.dbcsr_b200_engine_set_host_driver(self.h, ctypes.cast(self._host_driver_cb, _vp), None)

    def stats_cpu(self):
        """Totals of the stacks that took the host-driver route: dict(flop, entries, stacks)."""
        t = np.zeros(3, dtype=np.int64)
        self.L.dbcsr_b200_engine_stats_cpu(self.h, t.ctypes.data)
        return dict(flop=int(t[0]), entries=int(t[1]), stacks=int(t[2]))

    def finalize_c(self, filter_eps=None):
        """dbcsr_finalize on the device: optional final filter, every thread's blocks in BCSR order, data compacted."""
        rc = self.L.dbcsr_b200_engine_finalize_c(self.h, -1.0 if filter_eps is None else float(filter_eps))
        if rc != 0:
            raise acclib.AccError("dbcsr_b200_engine_finalize_c returned %d" % rc)

    def set_k_sizes(self, k_sizes):
        ks = np.ascontiguousarray(k_sizes, dtype=np.int32)
        if self.L.dbcsr_b200_engine_set_k_sizes(self.h, ks.ctypes.data, ks.size) != 0:
            raise acclib.AccError("dbcsr_b200_engine_set_k_sizes failed")

    def reset(self):
        if self.L.dbcsr_b200_engine_reset(self.h) != 0:
            raise acclib.AccError("dbcsr_b200_engine_reset failed")

    def sync(self):
        if self.L.dbcsr_b200_engine_sync(self.h) != 0:
            raise acclib.AccError("dbcsr_b200_engine_sync failed")

    @property
    def device_built_ticks(self):
        self.L.dbcsr_b200_engine_device_built_ticks.restype = ctypes.c_longlong
        self.L.dbcsr_b200_engine_device_built_ticks.argtypes = [_vp]
        return int(self.L.dbcsr_b200_engine_device_built_ticks(self.h))

    @property
    def nthreads(self):
        return self.L.dbcsr_b200_engine_nthreads(self.h)

    def c_index(self, t=0):
        nb = self.L.dbcsr_b200_engine_c_nblks(self.h, t)
        f = lambda fn: np.ctypeslib.as_array(fn(self.h, t), shape=(nb,)).copy() if nb else np.zeros(0, dtype=np.int32)
        return (f(self.L.dbcsr_b200_engine_c_rows), f(self.L.dbcsr_b200_engine_c_cols), f(self.L.dbcsr_b200_engine_c_blk_p),
                self.L.dbcsr_b200_engine_c_datasize(self.h, t))

    def c_dev(self, t=0):
        return self.L.dbcsr_b200_engine_c_dev(self.h, t)

    def c_to_host(self, t, host_array):
        if self.L.dbcsr_b200_engine_c_to_host(self.h, t, host_array.ctypes.data) != 0:
            raise acclib.AccError("dbcsr_b200_engine_c_to_host failed")

    def c_to_host_async(self, t, host_array):
        if self.L.dbcsr_b200_engine_c_to_host_async(self.h, t, host_array.ctypes.data) != 0:
            raise acclib.AccError("dbcsr_b200_engine_c_to_host_async failed")

    def set_c_host(self, t, host_array):
        ptr = host_array.ctypes.data if host_array is not None else None
        if self.L.dbcsr_b200_engine_set_c_host(self.h, t, ptr) != 0:
            raise acclib.AccError("dbcsr_b200_engine_set_c_host failed")

    def c_capacity(self, t):
        return int(self.L.dbcsr_b200_engine_c_capacity(self.h, t))

    def wait_event(self, event):
        if self.L.dbcsr_b200_engine_wait_event(self.h, event) != 0:
            raise acclib.AccError("dbcsr_b200_engine_wait_event failed")

    def flop(self):
        return int(self.L.dbcsr_b200_engine_flop(self.h))

    def build_seconds(self):
        return float(self.L.dbcsr_b200_engine_build_seconds(self.h))

    def stacks(self):
        """Recorded stacks: list of dicts like the index oracle produces (host S x 7, dev S x 3)."""
        out = []
        info = np.zeros(10, dtype=np.int32)
        for i in range(self.L.dbcsr_b200_engine_nstacks(self.h)):
            self.L.dbcsr_b200_engine_stack_info(self.h, i, info.ctypes.data)
            S = int(info[7])
            host = np.ctypeslib.as_array(self.L.dbcsr_b200_engine_stack_host(self.h, i), shape=(S, 7)).copy()
            dev = np.ctypeslib.as_array(self.L.dbcsr_b200_engine_stack_dev(self.h, i), shape=(S, 3)).copy()
            out.append(dict(m=int(info[0]), n=int(info[1]), k=int(info[2]), max_m=int(info[3]), max_n=int(info[4]), max_k=int(info[5]),
                            defined_mnk=bool(info[6]), thread=int(info[8]), stack_id=int(info[9]), host=host, dev=dev))
        return out

    def close(self):
        if self.h:
            self.L.dbcsr_b200_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Replay:
    """One rank's whole Cannon multiply on pre-built device stacks, enqueued by ONE C call per step (include/dbcsr_b200_host.h,
    dbcsr_b200_replay_*): peer pulls on a side stream, stack kernels on the compute stream, C zeroed once per multiply."""

    def __init__(self, nticks):
        self.L = _L()
        self.h = self.L.dbcsr_b200_replay_create(nticks)
        if not self.h:
            raise acclib.AccError("dbcsr_b200_replay_create failed")

    def _ck(self, rc, what):
        if rc != 0:
            raise acclib.AccError("dbcsr_b200_replay_%s returned %d" % (what, rc))

    def set_panels(self, tick, a_ptr, b_ptr):
        self._ck(self.L.dbcsr_b200_replay_set_panels(self.h, tick, a_ptr, b_ptr), "set_panels")

    def add_pull(self, tick, dst_ptr, src_ptr, nbytes):
        self._ck(self.L.dbcsr_b200_replay_add_pull(self.h, tick, dst_ptr, src_ptr, nbytes), "add_pull")

    def add_stack(self, tick, dev_stack_ptr, size, m, n, k, defined_mnk):
        self._ck(self.L.dbcsr_b200_replay_add_stack(self.h, tick, dev_stack_ptr, size, m, n, k, 1 if defined_mnk else 0), "add_stack")

    def set_c(self, c0_ptr, c1_ptr, nbytes, zero_overlap=True):
        self._ck(self.L.dbcsr_b200_replay_set_c(self.h, c0_ptr, c1_ptr, nbytes, 1 if zero_overlap else 0), "set_c")

    def step(self, compute_stream):
        self._ck(self.L.dbcsr_b200_replay_step(self.h, compute_stream), "step")

    def current_c(self):
        return self.L.dbcsr_b200_replay_current_c(self.h)

    def close(self):
        if self.h:
            self.L.dbcsr_b200_replay_destroy(self.h)
            self.h = None


def transpose_panel(acc, b_list3, k_sizes, n_sizes, b_dev_ptr, stream):
    """acc_transpose_blocks: transposes every block of the right panel in place on the device."""
    b = np.ascontiguousarray(b_list3, dtype=np.int32).reshape(-1, 3)
    ks = np.ascontiguousarray(k_sizes, dtype=np.int32)
    ns = np.ascontiguousarray(n_sizes, dtype=np.int32)
    nb = b.shape[0]
    scratch_h = acc.host_alloc((max(nb, 1),), np.int32)
    scratch_d = acc.dev_alloc(4 * max(nb, 1))
    rc = _L().dbcsr_b200_transpose_panel(b.ctypes.data, nb, ks.ctypes.data, ns.ctypes.data, b_dev_ptr, scratch_h.ptr, scratch_d.ptr, stream)
    acc.stream_sync(stream)
    scratch_h.free()
    scratch_d.free()
    if rc != 0:
        raise acclib.AccError("dbcsr_b200_transpose_panel returned %d" % rc)


def finalize_index(rows, cols, nelems):
    """dbcsr_finalize, index part (C++): returns (rows, cols, blk_p_new, perm, nze) of the BCSR-ordered index."""
    L = _L()
    r = np.ascontiguousarray(rows, dtype=np.int32).copy()
    c = np.ascontiguousarray(cols, dtype=np.int32).copy()
    ne = np.ascontiguousarray(nelems, dtype=np.int32)
    perm = np.empty(r.size, dtype=np.int32)
    bp = np.empty(r.size, dtype=np.int32)
    nze = ctypes.c_longlong(0)
    rc = L.dbcsr_b200_finalize_index(r.size, r.ctypes.data, c.ctypes.data, ne.ctypes.data, perm.ctypes.data, bp.ctypes.data, ctypes.byref(nze))
    if rc != 0:
        raise acclib.AccError("dbcsr_b200_finalize_index returned %d" % rc)
    return r, c, bp, perm, int(nze.value)
