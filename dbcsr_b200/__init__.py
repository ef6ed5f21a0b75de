"""dbcsr_b200 -- B200-native drop-in for DBCSR's accelerator layer (acc.h + acc_libsmm.h C ABI).

The product is the C-ABI shared library dbcsr_b200/lib/libdbcsr_acc_b200.so (sources in dbcsr_b200/csrc);
this Python package only binds it (dbcsr_b200.lib) and mirrors the host-side callers for benchmarks and tests.
"""
__version__ = "0.1.0"
