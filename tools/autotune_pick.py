"""Best kernel configuration per block size from a tools/kbench sweep (tools/autotune.sh): prints, per shape, the runs sorted by
TFLOP/s with their (warps per CTA, flush, align, chunk) decoded from the variant id, and flags parity failures."""
import re
import sys

WPC = {0: 2, 1: 4, 2: 8, 3: 12, 4: 16}
rows = {}
for line in open(sys.argv[1]):
    m = re.match(r"bsz\s+(\d+) spec (\S+)\s+variant\s+(\d+) balance (\d+) chunk\s+(\d+)\s+rc (-?\d+)\s+parity (\S+).*?mean ([\d.]+) ms.*?([\d.]+) TFLOP/s", line)
    if not m:
        continue
    bsz, spec, var, bal, chunk, rc, parity, ms, tf = m.groups()
    var, bal, chunk = int(var), int(bal), int(chunk)
    if 100 <= var < 150:
        desc = "warps/CTA %2d  flush %s" % (WPC[(var - 100) // 10], "bulk(stage)" if var % 10 == 2 else "RED")
    elif var in (154, 155, 158, 159):
        desc = "warps/CTA %2d  %d stages  flush RED" % (8 if var & 1 else 4, 2 if var < 158 else 4)
    else:
        desc = "variant %d" % var
    rows.setdefault(int(bsz), []).append((float(tf), desc, "aligned" if bal & 2 else "", "chunk %d" % chunk if chunk else "one wave", parity, int(rc)))
for bsz in sorted(rows):
    print("== %d^3" % bsz)
    for tf, desc, al, ch, parity, rc in sorted(rows[bsz], reverse=True)[:8]:
        print("  %7.2f TFLOP/s  %-34s %-8s %-10s %s%s" % (tf, desc, al, ch, parity, "" if rc == 0 else "  rc=%d" % rc))
    bad = [r for r in rows[bsz] if r[4] not in ("exact", "reference")]
    if bad:
        print("  PARITY FAILURES: %d" % len(bad))
