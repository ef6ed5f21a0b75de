#!/bin/bash
# Kernel-variant sweep on one GPU: runs the bench (value only) for every libvar_*.so next to the main library.
for lib in dbcsr_b200/lib/libdbcsr_acc_b200.so dbcsr_b200/lib/libvar_*.so; do
  [ -f "$lib" ] || continue
  echo "== $lib"
  DBCSR_B200_LIB=$PWD/$lib python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('value %.0f GF/s  kernel-only %.0f  avg launch %.2f us  frac %.3f' % (d['value'], r['kernel_only_gflops'], r['avg_launch_us'], r['frac']))
    elif 'Error' in l or 'error' in l: print(l.strip())
"
done
