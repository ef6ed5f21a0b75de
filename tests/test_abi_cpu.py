"""CPU-side checks of the drop-in boundary: the library loads without a GPU/driver and exports every symbol that
include/dbcsr_acc.h and include/dbcsr_acc_libsmm.h declare (no compute calls here)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

from dbcsr_b200 import lib as acclib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return set(re.findall(r"\b((?:c_dbcsr_|libsmm_acc_|c_calculate_)\w+)\s*\(", txt))


def test_library_exports_every_declared_symbol():
    L = acclib.load()
    declared = _declared_symbols("dbcsr_acc.h") | _declared_symbols("dbcsr_acc_libsmm.h")
    assert declared == set(acclib.ACC_SYMBOLS) | set(acclib.SMM_SYMBOLS), declared ^ (set(acclib.ACC_SYMBOLS) | set(acclib.SMM_SYMBOLS))
    for s in declared:
        assert hasattr(L, s), "missing export: " + s


def test_host_stand_in_header_symbols_are_exported():
    """include/dbcsr_b200_host.h (the C++ stand-in for DBCSR's Fortran local-multiply layer) is part of the same library."""
    txt = open(os.path.join(ROOT, "include", "dbcsr_b200_host.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    declared = set(re.findall(r"\b(dbcsr_b200_\w+)\s*\(", txt))
    assert len(declared) >= 30
    out = subprocess.check_output(["nm", "-D", "--defined-only", acclib.LIB_PATH]).decode()
    exported = set(l.split()[-1] for l in out.splitlines() if l.strip())
    assert not (declared - exported), declared - exported


def test_reference_abi_symbol_names_are_all_present():
    """The 26 + 6 names of the reference's src/acc/acc.h:34-71 and src/acc/acc_libsmm.h:38-49 (hard-coded list)."""
    ref = """c_dbcsr_acc_init c_dbcsr_acc_finalize c_dbcsr_acc_clear_errors c_dbcsr_acc_get_ndevices c_dbcsr_acc_set_active_device
    c_dbcsr_acc_device_synchronize c_dbcsr_acc_stream_priority_range c_dbcsr_acc_stream_create c_dbcsr_acc_stream_destroy
    c_dbcsr_acc_stream_sync c_dbcsr_acc_stream_wait_event c_dbcsr_acc_event_create c_dbcsr_acc_event_destroy c_dbcsr_acc_event_record
    c_dbcsr_acc_event_query c_dbcsr_acc_event_synchronize c_dbcsr_acc_dev_mem_allocate c_dbcsr_acc_dev_mem_deallocate
    c_dbcsr_acc_dev_mem_set_ptr c_dbcsr_acc_host_mem_allocate c_dbcsr_acc_host_mem_deallocate c_dbcsr_acc_memcpy_h2d
    c_dbcsr_acc_memcpy_d2h c_dbcsr_acc_memcpy_d2d c_dbcsr_acc_memset_zero c_dbcsr_acc_dev_mem_info
    libsmm_acc_init libsmm_acc_finalize libsmm_acc_is_thread_safe libsmm_acc_transpose libsmm_acc_process c_calculate_norms""".split()
    assert len(ref) == 32
    out = subprocess.check_output(["nm", "-D", "--defined-only", acclib.LIB_PATH]).decode()
    exported = set(l.split()[-1] for l in out.splitlines() if l.strip())
    assert not (set(ref) - exported), set(ref) - exported


def test_no_libcuda_link_dependency():
    """The library must load on a box without a driver (this container): only libc/libstdc++ are needed."""
    out = subprocess.check_output(["ldd", acclib.LIB_PATH]).decode()
    assert "libcuda.so" not in out and "not found" not in out


def test_thread_safety_flag_and_version():
    L = acclib.load()
    assert L.libsmm_acc_is_thread_safe() == 1  # src/core/dbcsr_lib.F:248-261 aborts otherwise in OpenMP builds
    assert L.libsmm_acc_gpu_warp_size() == 32
    assert b"sm_100a" in L.libsmm_acc_b200_version()
    assert L.libsmm_acc_b200_kernel_kind(23, 23, 23, 3) == 1
    assert L.libsmm_acc_b200_kernel_kind(7, 9, 11, 3) == 2
    assert L.libsmm_acc_b200_kernel_kind(23, 23, 23, 1) == 2  # real_4: typed generic kernel (reference: -10)
    assert L.libsmm_acc_b200_kernel_kind(23, 23, 23, 9) == 3  # BF16 extension
    assert L.libsmm_acc_b200_kernel_kind(23, 23, 23, 4) == 0


def test_product_does_not_reference_the_oracle():
    """The shipped path must never route through oracle/ (CPU) code: no import, no dlopen, no link."""
    pat = re.compile(r"^\s*(from|import)\s+oracle|liboracle|oracle\.(lib|ref)\(|oracle/_ref|orc_\w+\(", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dbcsr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not pat.search(txt), os.path.join(dirpath, f)
    out = subprocess.check_output(["nm", "-D", acclib.LIB_PATH]).decode()
    assert "orc_" not in out


def test_tunables_round_trip():
    """Run-time knobs of the FP64 stack kernels (dbcsr_b200/csrc/smm_tune.h): host-side state only, no device needed."""
    from dbcsr_b200 import lib as acclib

    L = acclib.load()
    assert L.libsmm_acc_b200_get_tunable(b"no such knob") == -1
    assert L.libsmm_acc_b200_set_tunable(b"no such knob", 1) == -1
    defaults = {}
    for name, val in (("balance", 1), ("align", 0), ("chunk", 7), ("trace_first", 3), ("trace_count", 2)):
        defaults[name] = L.libsmm_acc_b200_get_tunable(name.encode())
        assert L.libsmm_acc_b200_set_tunable(name.encode(), val) == 0
        assert L.libsmm_acc_b200_get_tunable(name.encode()) == val
    for name, val in defaults.items():
        assert L.libsmm_acc_b200_set_tunable(name.encode(), val) == 0
    # shipped defaults: per-shape launch policy (-1) for chunk and align, equal chunks off; production library has no variant table
    if "DBCSR_B200_CHUNK" not in os.environ and "DBCSR_B200_ALIGN" not in os.environ:
        assert defaults["chunk"] == -1 and defaults["align"] == -1
    assert L.libsmm_acc_b200_get_tunable(b"experiment") in (0, 1)


def test_autotune_database_and_generated_policy_are_in_sync():
    """dbcsr_b200/parameters/parameters_B200.json (one record per tuned (m,n,k), the B200 counterpart of the reference's
    parameters_<GPU>.json) and the generated smm_policy.inc the launch table includes."""
    import json

    db = json.load(open(os.path.join(ROOT, "dbcsr_b200", "parameters", "parameters_B200.json")))
    sizes = [5, 13, 23, 26, 32]
    assert sorted((r["m"], r["n"], r["k"]) for r in db) == sorted((m, n, k) for m in sizes for n in sizes for k in sizes)
    for r in db:
        assert r["algorithm"] in ("dmma", "tiny") and r["flush"] in (0, 2) and r["chunk"] >= 0 and r["source"].split(":")[0] in ("autotuned", "default")
        assert r["algorithm"] == "dmma" or r["m"] * r["n"] <= 96
        assert r.get("warps_per_cta", 0) in (0, 2, 4, 8, 12, 16)
        assert (r["perf"] > 0) == r["source"].startswith("autotuned")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_policy.py"), "--print"], capture_output=True, text=True, check=True).stdout
    assert out == open(os.path.join(ROOT, "dbcsr_b200", "csrc", "smm_policy.inc")).read()
    rec = [r for r in db if (r["m"], r["n"], r["k"]) == (23, 23, 23)][0]
    assert "SMM_POLICY(23, 23, 23, 0, %d, 2, %d, %s)" % (rec["warps_per_cta"], rec["chunk"], "true" if rec["align_runs"] else "false") in out


def test_reference_arm_runs_without_the_accelerator_library():
    """bench.py --impl reference times the oracle's CPU path on stacks from the host-only builder library: it must not need (or load)
    libdbcsr_acc_b200.so."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, DBCSR_B200_LIB="/nonexistent/libdbcsr_acc_b200.so")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-entries", "20000",
                          "--nblk", "80"], env=env, capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
