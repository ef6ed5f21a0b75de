"""
dbcsr_b200/multiply.py -- host-side mirror of the accelerator path of `dbcsr_multiply` for one rank and one Cannon tick
(C = A * B, alpha = 1, beta = 0, optional on-the-fly norm filter): what multiply_cannon + accdrv do around the C ABI
(src/mm/dbcsr_mm_cannon.F:1622-1667 host2dev of both panels + transpose of the right one; src/mm/dbcsr_mm_accdrv.F:340-362 D2H of C).

Everything that touches the device goes through the drop-in C ABI (dbcsr_b200.lib / dbcsr_b200.host); no CPU compute path.
"""
import numpy as np

from . import host
from . import lib as acclib


class ProductC:
    """Per-thread product work matrices: (rows, cols, blk_p (1-based), data) in order of first touch (pre-finalize)."""

    def __init__(self, m_sizes, n_sizes):
        self.m_sizes, self.n_sizes = np.asarray(m_sizes), np.asarray(n_sizes)
        self.parts = []

    def add(self, rows, cols, blk_p, data):
        self.parts.append((rows, cols, blk_p, data))

    @property
    def nblks(self):
        return sum(p[0].size for p in self.parts)

    def blocks(self):
        """dict (row, col) -> (m, n) ndarray"""
        out = {}
        for rows, cols, blk_p, data in self.parts:
            for r, c, o in zip(rows, cols, blk_p):
                m, n = int(self.m_sizes[r - 1]), int(self.n_sizes[c - 1])
                out[(int(r), int(c))] = data[o - 1:o - 1 + m * n].reshape(n, m).T
        return out

    def bcsr_index(self):
        """The finalized (canonical) index: row_p + sorted col_i (what dbcsr_finalize produces, independent of traversal)."""
        keys = sorted((int(r), int(c)) for rows, cols, _, _ in self.parts for r, c in zip(rows, cols))
        nrows = self.m_sizes.size
        row_p = np.zeros(nrows + 1, dtype=np.int64)
        for r, _ in keys:
            row_p[r] += 1
        return np.cumsum(row_p), np.array([c for _, c in keys], dtype=np.int32)


class DeviceMultiply:
    """Pooled resources for repeated multiplies of same-shaped panels (DBCSR keeps these in memory pools)."""

    def __init__(self, acc, m_sizes, n_sizes, k_sizes, a_nze_max, b_nze_max, nb_max, nthreads=1, cfg=None, c_capacity=0,
                 mode=host.LAUNCH):
        self.acc = acc
        self.m_sizes, self.n_sizes, self.k_sizes = (np.ascontiguousarray(x, dtype=np.int32) for x in (m_sizes, n_sizes, k_sizes))
        self.engine = host.Engine(self.m_sizes, self.n_sizes, self.k_sizes, nthreads=nthreads, mode=mode, cfg=cfg, c_capacity=c_capacity)
        self.copy_stream = acc.stream_create("panels", 0)
        self.d_a = acc.dev_alloc(8 * max(a_nze_max, 1))
        self.d_b = acc.dev_alloc(8 * max(b_nze_max, 1))
        self.trs_h = acc.host_alloc((2 * max(nb_max, 1),), np.int32)   # transpose stack + (fused norms) list positions
        self.trs_d = acc.dev_alloc(8 * max(nb_max, 1))
        self.d_bnorm = acc.dev_alloc(4 * max(nb_max, 1))
        self.b_norms_fused = None
        self.first = True
        self.panels_ready = acc.event_create()
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def upload_panels(self, a_data, b_data, b_list3, a_list3=None, want_b_norms=False):
        """host2dev of both panels + acc_transpose_blocks of the right one, all asynchronous on the copy stream; the event
        `panels_ready` orders the stack kernels behind them (the reference synchronises the whole device here,
        src/mm/dbcsr_mm_cannon.F:1642-1646; an event lets the host build stacks while the panels are still in flight).
        With a_list3 (BCSR-ordered left list, data offsets ascending in list order) the upload is pipelined: right panel and its
        transpose first (`panels_ready`), then the left panel in the engine's block-row chunks, one event per chunk - the stacks
        of the first rows start (and their C blocks travel back) while later rows are still going up."""
        acc = self.acc
        b = np.ascontiguousarray(b_list3, dtype=np.int32).reshape(-1, 3)
        pipelined = a_list3 is not None and self.engine.nchunks > 1
        if not pipelined:
            acc.h2d(a_data, self.d_a, self.copy_stream)
        acc.h2d(b_data, self.d_b, self.copy_stream)
        self.b_norms_fused = None
        rc = -3
        if want_b_norms and b.shape[0]:
            # transpose and squared block norms of the right panel in ONE pass (libsmm_acc_b200_transpose_norms); -3 = the panel has
            # blocks above max_kernel_dim, which are not transposed: separate passes as in the reference
            rc = acc.L.dbcsr_b200_transpose_panel_norms(b.ctypes.data, b.shape[0], self.k_sizes.ctypes.data, self.n_sizes.ctypes.data, self.d_b.ptr,
                                                        self.trs_h.ptr, self.trs_d.ptr, self.d_bnorm.ptr, self.copy_stream)
            if rc == 0:
                self.b_norms_fused = np.empty(b.shape[0], dtype=np.float32)
                acc.d2h(self.d_bnorm, self.b_norms_fused, self.copy_stream)
            elif rc != -3:
                raise acclib.AccError("transpose_panel_norms returned %d" % rc)
        if rc == -3:
            rc = acc.L.dbcsr_b200_transpose_panel(b.ctypes.data, b.shape[0], self.k_sizes.ctypes.data, self.n_sizes.ctypes.data, self.d_b.ptr,
                                                  self.trs_h.ptr, self.trs_d.ptr, self.copy_stream)
            if rc != 0:
                raise acclib.AccError("transpose_panel returned %d" % rc)
        acc.event_record(self.panels_ready, self.copy_stream)
        self._chunk_events_armed = None
        if pipelined:
            a = np.ascontiguousarray(a_list3, dtype=np.int32).reshape(-1, 3)
            nch = self.engine.nchunks
            if getattr(self, "_chunk_events", None) is None or len(self._chunk_events) != nch:
                self._chunk_events = [acc.event_create() for _ in range(nch)]
            # element range of every chunk in the left data area: blocks of rows (row_lo, row_hi] are contiguous in BCSR order
            row_hi = np.array([self.engine.chunk_rows(c)[1] for c in range(nch)], dtype=np.int64)
            blk_end = np.searchsorted(a[:, 0], row_hi, side="right")  # first list position beyond the chunk
            ends = np.where(blk_end < a.shape[0], a[np.minimum(blk_end, a.shape[0] - 1), 2].astype(np.int64) - 1, a_data.size) \
                if a.shape[0] else np.zeros(nch, dtype=np.int64)
            ends[-1] = a_data.size
            lo = 0
            for c in range(nch):
                hi = int(ends[c])
                if hi > lo:
                    acc.h2d(a_data[lo:hi], self.d_a, self.copy_stream, offset_bytes=8 * lo)
                    lo = hi
                acc.event_record(self._chunk_events[c], self.copy_stream)
            self._chunk_events_armed = list(self._chunk_events)
        self._keep = (a_data, b_data, b)  # host buffers must stay alive until the copies ran
        self.h2d_bytes = a_data.nbytes + b_data.nbytes + 4 * b.shape[0]

    def panel_norms(self, list3, row_sizes, col_sizes, d_data):
        """acc_calculate_norms (src/mm/dbcsr_mm_common.F:498-591): offsets (blk_p - 1) and element counts go up, c_calculate_norms
        runs on the device data area, the squared single-precision block norms come back.  Runs on the copy stream behind the
        panel upload; the in-place transpose of the right panel does not change a block's sum of squares."""
        acc = self.acc
        l3 = np.ascontiguousarray(list3, dtype=np.int32).reshape(-1, 3)
        nb = l3.shape[0]
        if nb == 0:
            return np.zeros(0, dtype=np.float32)
        offs = np.ascontiguousarray(l3[:, 2] - 1, dtype=np.int32)
        nel = np.ascontiguousarray(np.asarray(row_sizes)[l3[:, 0] - 1] * np.asarray(col_sizes)[l3[:, 1] - 1], dtype=np.int32)
        d_o, d_n, d_out = acc.dev_alloc(4 * nb), acc.dev_alloc(4 * nb), acc.dev_alloc(4 * nb)
        out = np.empty(nb, dtype=np.float32)
        try:
            acc.h2d(offs, d_o, self.copy_stream)
            acc.h2d(nel, d_n, self.copy_stream)
            acc.norms(d_data.ptr, nb, d_o.ptr, d_n.ptr, d_out.ptr, self.copy_stream)
            acc.d2h(d_out, out, self.copy_stream)
            acc.stream_sync(self.copy_stream)
        finally:
            for d in (d_o, d_n, d_out):
                d.free()
        return out

    def multiply(self, a_list3, b_list3, filter_eps=None, total_row_counts=None, c_preset=None, retain_sparsity=False,
                 c_symmetry=False):
        """One local multiply on the uploaded panels (stacks are built, ordered, uploaded and drained asynchronously).
        filter_eps: dbcsr_multiply's on-the-fly filter; total_row_counts = A blocks per block row over the whole process row
        (default: of this panel, i.e. a 1-column process grid).
        c_preset = (rows, cols, data): existing C blocks, data already scaled by beta (C = A*B + beta*C_old); with
        retain_sparsity only products landing in those blocks are computed.
        c_symmetry: the product has symmetry - the mirrored half of the off-diagonal blocks is not computed
        (checkerboard rule, src/mm/dbcsr_mm_csr.F:280-292)."""
        if not self.first:
            self.engine.reset()
        self.first = False
        if c_preset is not None:
            self.engine.preset_c(c_preset[0], c_preset[1], c_preset[2], keep_sparsity=retain_sparsity)
            self.h2d_bytes += 0 if c_preset[2] is None else 8 * int(np.asarray(c_preset[2]).size)
        if c_symmetry:
            self.engine.set_c_symmetry(True)
        self.engine.wait_event(self.panels_ready)
        if getattr(self, "_chunk_events_armed", None):
            self.engine.set_chunk_events(self._chunk_events_armed)
            self._chunk_events_armed = None
        if filter_eps is None:
            self.engine.set_filter(None)
            self.engine.multiply(a_list3, self.d_a.ptr, b_list3, self.d_b.ptr)
            return
        a = np.ascontiguousarray(a_list3, dtype=np.int32).reshape(-1, 3)
        if total_row_counts is None:
            total_row_counts = np.bincount(a[:, 0] - 1, minlength=self.m_sizes.size)
        self.a_norms = self.panel_norms(a, self.m_sizes, self.k_sizes, self.d_a)
        if self.b_norms_fused is not None:  # computed together with the transpose (upload_panels(want_b_norms=True))
            acc_ = self.acc
            acc_.stream_sync(self.copy_stream)
            self.b_norms = self.b_norms_fused
        else:
            self.b_norms = self.panel_norms(b_list3, self.k_sizes, self.n_sizes, self.d_b)
        self.engine.set_filter(host.row_max_epss(filter_eps, total_row_counts))
        self.engine.multiply(a, self.d_a.ptr, b_list3, self.d_b.ptr, a_norms=self.a_norms, b_norms=self.b_norms)

    def filter_c(self, filter_eps):
        """Final filter of the product on the device before the download (multrec_filtering, src/mm/dbcsr_mm_multrec.F:700-758):
        blocks with sum(x^2) < filter_eps^2 are dropped, the survivors are packed contiguously (less PCIe traffic)."""
        assert getattr(self, "early", None) is None, "early per-thread D2H and the final filter exclude each other"
        self.engine.filter_c(filter_eps)

    def finalize_c(self, filter_eps=None):
        """dbcsr_finalize on the device before the download (work/dbcsr_work_operations.F:749-958 + the final filter when
        filter_eps is given): every thread's blocks in BCSR order, data compacted - what comes over PCIe is final."""
        assert getattr(self, "early", None) is None, "early per-thread D2H and the device finalize exclude each other"
        self.engine.finalize_c(filter_eps)

    def set_result_buffers(self, out_arrays):
        """Pooled (pinned) host buffers for C, one per thread, each at least engine.c_capacity(t) elements: from now on every
        thread enqueues its own D2H right behind its last stack (download_c then only waits)."""
        self.early = list(out_arrays)
        for t, arr in enumerate(self.early):
            assert arr.size >= self.engine.c_capacity(t)
            self.engine.set_c_host(t, arr)

    def download_c(self, out_arrays=None):
        """D2H of every thread's C buffer (datasize elements each). out_arrays: optional list of (pinned) host arrays."""
        prod = ProductC(self.m_sizes, self.n_sizes)
        self.d2h_bytes = 0
        if getattr(self, "early", None) is not None:  # copies were enqueued by the engine threads themselves
            self.engine.sync()
            for t in range(self.engine.nthreads):
                rows, cols, blk_p, ds = self.engine.c_index(t)
                self.d2h_bytes += 8 * ds
                prod.add(rows, cols, blk_p, self.early[t][:ds])
            return prod
        # every thread's D2H is enqueued behind that thread's last stack, so early finishers copy while others still compute
        for t in range(self.engine.nthreads):
            rows, cols, blk_p, ds = self.engine.c_index(t)
            buf = out_arrays[t][:ds] if out_arrays is not None else np.empty(ds)
            if ds:
                self.engine.c_to_host_async(t, buf)
            self.d2h_bytes += 8 * ds
            prod.add(rows, cols, blk_p, buf)
        self.engine.sync()
        return prod

    def close(self):
        self.engine.close()
        for d in (self.d_a, self.d_b, self.trs_d, self.d_bnorm):
            d.free()
        self.trs_h.free()
        self.acc.event_destroy(self.panels_ready)
        for ev in getattr(self, "_chunk_events", None) or []:
            self.acc.event_destroy(ev)
        self.acc.stream_destroy(self.copy_stream)


def multiply(acc, A, B, m_sizes, n_sizes, k_sizes, nthreads=1, cfg=None, pipelined=False):
    """Convenience one-shot: C = A*B for host panels A, B (dbcsr_b200.workload.Panel-like: .data, .list3())."""
    dm = DeviceMultiply(acc, m_sizes, n_sizes, k_sizes, A.data.size, B.data.size, B.nblks, nthreads=nthreads, cfg=cfg)
    try:
        dm.upload_panels(A.data, B.data, B.list3(), a_list3=A.list3() if pipelined else None)
        dm.multiply(A.list3(), B.list3())
        return dm.download_c(), dm.engine.flop()
    finally:
        dm.close()
