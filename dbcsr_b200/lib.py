"""
dbcsr_b200/lib.py -- ctypes binding of the drop-in C ABI (include/dbcsr_acc.h, include/dbcsr_acc_libsmm.h).

This is the same binding surface DBCSR's Fortran layer uses through ISO_C_BINDING (src/acc/dbcsr_acc_*.F,
src/mm/dbcsr_acc_operations.F:38-67): opaque stream/event handles, raw device pointers, 1-based stacks.  There is NO
fall-back: if dbcsr_b200/lib/libdbcsr_acc_b200.so is missing or a call fails, an exception is raised.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# DBCSR_B200_LIB selects an alternative build of the same library (kernel-variant experiments); default = the in-tree build
LIB_PATH = os.environ.get("DBCSR_B200_LIB") or os.path.join(_HERE, "lib", "libdbcsr_acc_b200.so")

# every symbol include/*.h declares; tests check that the library exports all of them
ACC_SYMBOLS = [
    "c_dbcsr_acc_init", "c_dbcsr_acc_finalize", "c_dbcsr_acc_clear_errors", "c_dbcsr_acc_get_ndevices",
    "c_dbcsr_acc_set_active_device", "c_dbcsr_acc_device_synchronize", "c_dbcsr_acc_stream_priority_range",
    "c_dbcsr_acc_stream_create", "c_dbcsr_acc_stream_destroy", "c_dbcsr_acc_stream_sync", "c_dbcsr_acc_stream_wait_event",
    "c_dbcsr_acc_event_create", "c_dbcsr_acc_event_destroy", "c_dbcsr_acc_event_record", "c_dbcsr_acc_event_query",
    "c_dbcsr_acc_event_synchronize", "c_dbcsr_acc_dev_mem_allocate", "c_dbcsr_acc_dev_mem_deallocate",
    "c_dbcsr_acc_dev_mem_set_ptr", "c_dbcsr_acc_host_mem_allocate", "c_dbcsr_acc_host_mem_deallocate",
    "c_dbcsr_acc_memcpy_h2d", "c_dbcsr_acc_memcpy_d2h", "c_dbcsr_acc_memcpy_d2d", "c_dbcsr_acc_memset_zero",
    "c_dbcsr_acc_dev_mem_info", "c_dbcsr_timeset", "c_dbcsr_timestop",
]
SMM_SYMBOLS = [
    "libsmm_acc_init", "libsmm_acc_finalize", "libsmm_acc_is_thread_safe", "libsmm_acc_transpose", "libsmm_acc_process",
    "c_calculate_norms", "libsmm_acc_gpu_warp_size", "libsmm_acc_b200_kernel_kind", "libsmm_acc_b200_launch_count",
    "libsmm_acc_b200_version", "libsmm_acc_b200_pack_bf16", "libsmm_acc_b200_bf16_tile_bytes",
    "libsmm_acc_b200_block_norms_f64", "libsmm_acc_b200_gather_blocks", "libsmm_acc_b200_set_tunable",
    "libsmm_acc_b200_get_tunable", "libsmm_acc_b200_set_trace", "libsmm_acc_b200_stream_chain", "libsmm_acc_b200_fp64_peak_gflops", "libsmm_acc_b200_fp64_peak_sustained_gflops",
    "libsmm_acc_b200_memset_zero_trickle", "libsmm_acc_b200_transpose_norms",
    "libsmm_acc_b200_bf16_rk_tile_bytes", "libsmm_acc_b200_bf16_rk_slot_bytes", "libsmm_acc_b200_pack_bf16_rk", "libsmm_acc_b200_bf16_spgemm",
]

DBCSR_TYPE_REAL_8 = 3
DBCSR_TYPE_BF16_EXT = 9  # extension of this library: BF16 operand tiles, FP32 C
MAX_KERNEL_DIM = 80  # src/core/dbcsr_config.F:185

_vp, _i, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
_lib = None


class AccError(RuntimeError):
    pass


def load():
    """Load the C-ABI library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AccError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "or `make -C dbcsr_b200/csrc`" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    L.c_dbcsr_acc_get_ndevices.argtypes = [ctypes.POINTER(_i)]
    L.c_dbcsr_acc_set_active_device.argtypes = [_i]
    L.c_dbcsr_acc_stream_priority_range.argtypes = [ctypes.POINTER(_i), ctypes.POINTER(_i)]
    L.c_dbcsr_acc_stream_create.argtypes = [ctypes.POINTER(_vp), ctypes.c_char_p, _i]
    L.c_dbcsr_acc_stream_destroy.argtypes = [_vp]
    L.c_dbcsr_acc_stream_sync.argtypes = [_vp]
    L.c_dbcsr_acc_stream_wait_event.argtypes = [_vp, _vp]
    L.c_dbcsr_acc_event_create.argtypes = [ctypes.POINTER(_vp)]
    L.c_dbcsr_acc_event_destroy.argtypes = [_vp]
    L.c_dbcsr_acc_event_record.argtypes = [_vp, _vp]
    L.c_dbcsr_acc_event_query.argtypes = [_vp, ctypes.POINTER(_i)]
    L.c_dbcsr_acc_event_synchronize.argtypes = [_vp]
    L.c_dbcsr_acc_dev_mem_allocate.argtypes = [ctypes.POINTER(_vp), _sz]
    L.c_dbcsr_acc_dev_mem_deallocate.argtypes = [_vp]
    L.c_dbcsr_acc_dev_mem_set_ptr.argtypes = [ctypes.POINTER(_vp), _vp, _sz]
    L.c_dbcsr_acc_host_mem_allocate.argtypes = [ctypes.POINTER(_vp), _sz, _vp]
    L.c_dbcsr_acc_host_mem_deallocate.argtypes = [_vp, _vp]
    L.c_dbcsr_acc_memcpy_h2d.argtypes = [_vp, _vp, _sz, _vp]
    L.c_dbcsr_acc_memcpy_d2h.argtypes = [_vp, _vp, _sz, _vp]
    L.c_dbcsr_acc_memcpy_d2d.argtypes = [_vp, _vp, _sz, _vp]
    L.c_dbcsr_acc_memset_zero.argtypes = [_vp, _sz, _sz, _vp]
    L.c_dbcsr_acc_dev_mem_info.argtypes = [ctypes.POINTER(_sz), ctypes.POINTER(_sz)]
    L.libsmm_acc_transpose.argtypes = [_vp, _i, _i, _vp, _i, _i, _i, _i, _vp]
    L.libsmm_acc_process.argtypes = [_vp, _vp, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]
    L.c_calculate_norms.argtypes = [_vp, _i, _vp, _vp, _vp, _vp]
    L.libsmm_acc_b200_block_norms_f64.argtypes = [_vp, _i, _vp, _vp, _vp, _vp]
    L.libsmm_acc_b200_gather_blocks.argtypes = [_vp, _vp, _i, _vp, _vp, _vp, _vp]
    L.libsmm_acc_b200_kernel_kind.argtypes = [_i, _i, _i, _i]
    L.libsmm_acc_b200_launch_count.restype = ctypes.c_longlong
    L.libsmm_acc_b200_pack_bf16.argtypes = [_vp, _i, _i, _i, _i, _i, _vp, _vp]
    L.libsmm_acc_b200_bf16_tile_bytes.argtypes = [_i, _i]
    L.libsmm_acc_b200_version.restype = ctypes.c_char_p
    L.libsmm_acc_b200_set_tunable.argtypes = [ctypes.c_char_p, ctypes.c_longlong]
    L.libsmm_acc_b200_get_tunable.argtypes = [ctypes.c_char_p]
    L.libsmm_acc_b200_get_tunable.restype = ctypes.c_longlong
    L.libsmm_acc_b200_set_trace.argtypes = [_vp]
    L.libsmm_acc_b200_set_trace.restype = None
    L.libsmm_acc_b200_stream_chain.argtypes = [_vp, _i]
    L.libsmm_acc_b200_bf16_rk_tile_bytes.argtypes = [_i]
    L.libsmm_acc_b200_bf16_rk_slot_bytes.argtypes = [_i, _i]
    L.libsmm_acc_b200_pack_bf16_rk.argtypes = [_vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp]
    L.libsmm_acc_b200_bf16_spgemm.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]
    L.libsmm_acc_b200_memset_zero_trickle.argtypes = [_vp, _sz, _sz, _i, _vp]
    L.libsmm_acc_b200_transpose_norms.argtypes = [_vp, _vp, _i, _i, _vp, _i, _i, _i, _vp, _vp]
    L.libsmm_acc_b200_fp64_peak_sustained_gflops.argtypes = [_vp, ctypes.c_double]
    L.libsmm_acc_b200_fp64_peak_sustained_gflops.restype = ctypes.c_double
    L.libsmm_acc_b200_fp64_peak_gflops.argtypes = [_vp]
    L.libsmm_acc_b200_fp64_peak_gflops.restype = ctypes.c_double
    L.c_dbcsr_acc_clear_errors.restype = None
    _lib = L
    return L


def _ck(rc, what):
    if rc != 0:
        raise AccError("%s returned %d" % (what, rc))


class DevMem:
    """A device allocation obtained through c_dbcsr_acc_dev_mem_allocate (raw pointer owned by the caller)."""

    def __init__(self, acc, nbytes):
        self.acc, self.nbytes = acc, int(nbytes)
        p = _vp()
        _ck(acc.L.c_dbcsr_acc_dev_mem_allocate(ctypes.byref(p), max(self.nbytes, 1)), "dev_mem_allocate")
        self.ptr = p.value

    def free(self):
        if self.ptr:
            _ck(self.acc.L.c_dbcsr_acc_dev_mem_deallocate(self.ptr), "dev_mem_deallocate")
            self.ptr = None


class HostMem:
    """Pinned host buffer from c_dbcsr_acc_host_mem_allocate, exposed as a numpy array."""

    def __init__(self, acc, shape, dtype):
        self.acc = acc
        dtype = np.dtype(dtype)
        n = int(np.prod(shape))
        self.nbytes = n * dtype.itemsize
        p = _vp()
        _ck(acc.L.c_dbcsr_acc_host_mem_allocate(ctypes.byref(p), max(self.nbytes, 1), None), "host_mem_allocate")
        self.ptr = p.value
        buf = (ctypes.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            _ck(self.acc.L.c_dbcsr_acc_host_mem_deallocate(self.ptr, None), "host_mem_deallocate")
            self.ptr = None


class Acc:
    """Object-style view of the C ABI, mirroring DBCSR's Fortran wrappers (acc_stream_type, acc_event_type, acc_devmem_type)."""

    def __init__(self, device=0):
        self.L = load()
        n = _i(0)
        _ck(self.L.c_dbcsr_acc_get_ndevices(ctypes.byref(n)), "get_ndevices")
        if n.value <= 0:
            raise AccError("no CUDA device visible: the B200 path cannot run (and there is no CPU fall-back)")
        self.ndevices = n.value
        _ck(self.L.c_dbcsr_acc_set_active_device(device), "set_active_device")
        _ck(self.L.c_dbcsr_acc_init(), "init")
        self.device = device

    # streams / events ---------------------------------------------------------------------------------------------
    def stream_create(self, name="stream", priority=0):
        s = _vp()
        _ck(self.L.c_dbcsr_acc_stream_create(ctypes.byref(s), name.encode(), priority), "stream_create")
        return s.value

    def stream_destroy(self, s):
        _ck(self.L.c_dbcsr_acc_stream_destroy(s), "stream_destroy")

    def stream_sync(self, s):
        _ck(self.L.c_dbcsr_acc_stream_sync(s), "stream_sync")

    def event_create(self):
        e = _vp()
        _ck(self.L.c_dbcsr_acc_event_create(ctypes.byref(e)), "event_create")
        return e.value

    def event_record(self, e, s):
        _ck(self.L.c_dbcsr_acc_event_record(e, s), "event_record")

    def event_destroy(self, e):
        _ck(self.L.c_dbcsr_acc_event_destroy(e), "event_destroy")

    def stream_wait_event(self, s, e):
        _ck(self.L.c_dbcsr_acc_stream_wait_event(s, e), "stream_wait_event")

    def device_synchronize(self):
        _ck(self.L.c_dbcsr_acc_device_synchronize(), "device_synchronize")

    # memory -------------------------------------------------------------------------------------------------------
    def dev_alloc(self, nbytes):
        return DevMem(self, nbytes)

    def host_alloc(self, shape, dtype):
        return HostMem(self, shape, dtype)

    def h2d(self, host_array, dev, stream, offset_bytes=0):
        a = np.ascontiguousarray(host_array)
        _ck(self.L.c_dbcsr_acc_memcpy_h2d(a.ctypes.data, dev.ptr + offset_bytes, a.nbytes, stream), "memcpy_h2d")
        return a  # keep alive until the stream is synchronised

    def d2h(self, dev, host_array, stream, offset_bytes=0):
        assert host_array.flags["C_CONTIGUOUS"]
        _ck(self.L.c_dbcsr_acc_memcpy_d2h(dev.ptr + offset_bytes, host_array.ctypes.data, host_array.nbytes, stream), "memcpy_d2h")

    def memset_zero(self, dev, stream, offset=0, nbytes=None):
        _ck(self.L.c_dbcsr_acc_memset_zero(dev.ptr, offset, dev.nbytes - offset if nbytes is None else nbytes, stream), "memset_zero")

    def to_device(self, host_array, stream):
        """Allocate + upload + sync (convenience for tests)."""
        a = np.ascontiguousarray(host_array)
        d = self.dev_alloc(a.nbytes)
        if a.nbytes:
            self.h2d(a, d, stream)
            self.stream_sync(stream)
        return d

    def to_host(self, dev, shape, dtype, stream):
        out = np.empty(shape, dtype=dtype)
        if out.nbytes:
            self.d2h(dev, out, stream)
            self.stream_sync(stream)
        return out

    # the hot path ---------------------------------------------------------------------------------------------------
    def process(self, host_stack7, dev_stack3_ptr, stack_size, a_ptr, b_ptr, c_ptr, m, n, k, def_mnk, stack_stream, c_stream,
                datatype=DBCSR_TYPE_REAL_8, max_kernel_dim=MAX_KERNEL_DIM):
        """libsmm_acc_process: returns the reference's code (0, 10, or <0 = not run, C untouched)."""
        hp = host_stack7.ctypes.data if host_stack7 is not None else None
        return self.L.libsmm_acc_process(hp, dev_stack3_ptr, stack_size, datatype, a_ptr, b_ptr, c_ptr, m, n, k, max_kernel_dim,
                                         1 if def_mnk else 0, stack_stream, c_stream)

    def transpose(self, dev_trs_stack_ptr, offset, nblks, data_ptr, m, n, stream, datatype=DBCSR_TYPE_REAL_8,
                  max_kernel_dim=MAX_KERNEL_DIM):
        _ck(self.L.libsmm_acc_transpose(dev_trs_stack_ptr, offset, nblks, data_ptr, datatype, m, n, max_kernel_dim, stream),
            "libsmm_acc_transpose")

    def norms(self, mat_ptr, nblks, offsets_ptr, nelems_ptr, norms_ptr, stream):
        _ck(self.L.c_calculate_norms(mat_ptr, nblks, offsets_ptr, nelems_ptr, norms_ptr, stream), "c_calculate_norms")

    def block_norms_f64(self, mat_ptr, nblks, offsets_ptr, nelems_ptr, norms_ptr, stream):
        _ck(self.L.libsmm_acc_b200_block_norms_f64(mat_ptr, nblks, offsets_ptr, nelems_ptr, norms_ptr, stream), "block_norms_f64")

    def gather_blocks(self, src_ptr, dst_ptr, nblks, src_off_ptr, dst_off_ptr, nelems_ptr, stream):
        _ck(self.L.libsmm_acc_b200_gather_blocks(src_ptr, dst_ptr, nblks, src_off_ptr, dst_off_ptr, nelems_ptr, stream), "gather_blocks")

    def pack_bf16(self, src_ptr, nblks, rows, kdim, row_stride, k_stride, dst_ptr, stream):
        _ck(self.L.libsmm_acc_b200_pack_bf16(src_ptr, nblks, rows, kdim, row_stride, k_stride, dst_ptr, stream), "libsmm_acc_b200_pack_bf16")

    def bf16_tile_bytes(self, rows, kdim):
        return int(self.L.libsmm_acc_b200_bf16_tile_bytes(rows, kdim))

    def stream_chain(self, stream, on=True):
        """Declare `stream` a chain of independent stack drains (programmatic dependent launch without the grid-dependency wait in
        front of the reads, include/dbcsr_acc_libsmm.h); on=False withdraws the declaration."""
        _ck(self.L.libsmm_acc_b200_stream_chain(stream, 1 if on else 0), "stream_chain")

    def bf16_rk_tile_bytes(self, rows):
        return int(self.L.libsmm_acc_b200_bf16_rk_tile_bytes(rows))

    def bf16_rk_slot_bytes(self, rows, b_operand):
        return int(self.L.libsmm_acc_b200_bf16_rk_slot_bytes(rows, 1 if b_operand else 0))

    def pack_bf16_rk(self, src_ptr, nblks, rows, kdim, row_stride, k_stride, dst_ptr, dst_pitch, dst_slot_ptr, stream):
        _ck(self.L.libsmm_acc_b200_pack_bf16_rk(src_ptr, nblks, rows, kdim, row_stride, k_stride, dst_ptr, dst_pitch, dst_slot_ptr, stream),
            "pack_bf16_rk")

    def bf16_spgemm(self, a_tiles_ptr, a_map_ptr, b_tiles_ptr, b_map_ptr, c_ptr, c_off_ptr, nrb, ncb, nkb, m, n, k, stream):
        _ck(self.L.libsmm_acc_b200_bf16_spgemm(a_tiles_ptr, a_map_ptr, b_tiles_ptr, b_map_ptr, c_ptr, c_off_ptr, nrb, ncb, nkb, m, n, k, stream),
            "bf16_spgemm")

    def memset_zero_trickle(self, dev, stream, nctas=8, nbytes=None):
        """Zero `dev` with `nctas` CTAs only (bounded rate), see include/dbcsr_acc_libsmm.h."""
        n = dev.nbytes if nbytes is None else nbytes
        _ck(self.L.libsmm_acc_b200_memset_zero_trickle(dev.ptr, 0, n // 16 * 16, nctas, stream), "memset_zero_trickle")
        if n % 16:
            _ck(self.L.c_dbcsr_acc_memset_zero(dev.ptr, n // 16 * 16, n % 16, stream), "memset_zero")

    def fp64_peak_gflops(self, stream):
        """Measured DMMA.8x8x4 throughput of this device (register operands), GFLOP/s."""
        return float(self.L.libsmm_acc_b200_fp64_peak_gflops(stream))

    def fp64_peak_sustained_gflops(self, stream, seconds=0.4):
        """The same loop back to back for `seconds`: throughput over the second half (sustained, under the power limit)."""
        return float(self.L.libsmm_acc_b200_fp64_peak_sustained_gflops(stream, float(seconds)))

    def launch_count(self):
        return int(self.L.libsmm_acc_b200_launch_count())

    def set_tunable(self, name, value):
        """Run-time knob of the FP64 stack kernels (include/dbcsr_acc_libsmm.h): "balance", "align", "chunk", "variant", ..."""
        _ck(self.L.libsmm_acc_b200_set_tunable(name.encode(), int(value)), "libsmm_acc_b200_set_tunable(%s)" % name)

    def get_tunable(self, name):
        return int(self.L.libsmm_acc_b200_get_tunable(name.encode()))

    def finalize(self):
        _ck(self.L.c_dbcsr_acc_finalize(), "finalize")
