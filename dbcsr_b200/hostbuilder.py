"""
dbcsr_b200/hostbuilder.py -- ctypes binding of libdbcsr_b200_hostbuilder.so (csrc/host/record_engine.cpp): DBCSR's host-side stack
building (rec_sort_index, sparse_multrec, csr_multiply_low, flush_stacks; one row slice per thread) WITHOUT any accelerator code.
Used by bench.py's CPU reference arm, which must not depend on the accelerator library.
"""
import ctypes
import os

import numpy as np

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libdbcsr_b200_hostbuilder.so")
        if not os.path.exists(path):
            raise RuntimeError("%s is missing: run python -c 'import __graft_entry__ as g; g.build()'" % path)
        L = ctypes.CDLL(path)
        vp, i = ctypes.c_void_p, ctypes.c_int
        L.dbcsr_b200_recorder_run.argtypes = [vp, i, vp, i, vp, i, vp, i, vp, i, i, i, i]
        L.dbcsr_b200_recorder_run.restype = vp
        L.dbcsr_b200_recorder_nstacks.argtypes = [vp]
        L.dbcsr_b200_recorder_stack_info.argtypes = [vp, i, vp]
        L.dbcsr_b200_recorder_stack_host.argtypes = [vp, i]
        L.dbcsr_b200_recorder_stack_host.restype = ctypes.POINTER(ctypes.c_int)
        L.dbcsr_b200_recorder_datasize.argtypes = [vp, i]
        L.dbcsr_b200_recorder_flop.argtypes = [vp]
        L.dbcsr_b200_recorder_flop.restype = ctypes.c_longlong
        L.dbcsr_b200_recorder_free.argtypes = [vp]
        _LIB = L
    return _LIB


def record_stacks(m_sizes, n_sizes, k_sizes, a_list3, b_list3, nthreads=1, mm_stack_size=1000, n_stacks=3):
    """Builds the stacks of C = A * B the way `nthreads` DBCSR threads would (thread t owns block rows (t*nrows/T, (t+1)*nrows/T]).
    Returns (stacks, datasizes, flop): stacks = list of dicts(m, n, k, defined_mnk, thread, stack_id, host = S x 7 int32) in dispatch
    order thread by thread; datasizes[t] = elements of thread t's C work area."""
    L = _lib()
    ms, ns, ks = (np.ascontiguousarray(x, dtype=np.int32) for x in (m_sizes, n_sizes, k_sizes))
    a = np.ascontiguousarray(a_list3, dtype=np.int32).reshape(-1, 3)
    b = np.ascontiguousarray(b_list3, dtype=np.int32).reshape(-1, 3)
    h = L.dbcsr_b200_recorder_run(ms.ctypes.data, ms.size, ns.ctypes.data, ns.size, ks.ctypes.data, ks.size, a.ctypes.data, a.shape[0],
                                  b.ctypes.data, b.shape[0], int(nthreads), int(mm_stack_size), int(n_stacks))
    if not h:
        raise RuntimeError("dbcsr_b200_recorder_run failed")
    try:
        info = np.zeros(7, dtype=np.int32)
        stacks = []
        for i in range(L.dbcsr_b200_recorder_nstacks(h)):
            L.dbcsr_b200_recorder_stack_info(h, i, info.ctypes.data)
            S = int(info[4])
            host = np.ctypeslib.as_array(L.dbcsr_b200_recorder_stack_host(h, i), shape=(S, 7)).copy()
            stacks.append(dict(m=int(info[0]), n=int(info[1]), k=int(info[2]), defined_mnk=bool(info[3]), thread=int(info[5]), stack_id=int(info[6]),
                               host=host))
        datasizes = [int(L.dbcsr_b200_recorder_datasize(h, t)) for t in range(int(nthreads))]
        return stacks, datasizes, int(L.dbcsr_b200_recorder_flop(h))
    finally:
        L.dbcsr_b200_recorder_free(h)
