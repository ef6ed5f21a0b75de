"""Writes the autotune database dbcsr_b200/parameters/parameters_B200.json from a tools/autotune_all.sh sweep
(gpurun_out/kbench_results.txt or a copy under profiles/): per (m,n,k) the fastest parity-exact configuration.
  python tools/autotune_db.py <results.txt> [--source "text for the source field"] [--dry]
Variant ids (dbcsr_b200/csrc/smm_inst.cu): 9 = the library default for the shape, 100 + 10*i + f = DMMA kernel with warps per CTA
{2,4,8,12,16}[i] and flush f (0 RED, 2 TMA bulk reduction), 200 / 201 = lane-per-element kernel with 8 / 4 warps per CTA.
kbench's balance field: bit 1 = run-aligned chunks; chunk = entries per warp (0 = one resident wave)."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WPC = {0: 2, 1: 4, 2: 8, 3: 12, 4: 16}
LINE = re.compile(r"bsz\s+\d+ mnk (\d+),(\d+),(\d+) spec (\S+)\s+variant\s+(\d+) balance (-?\d+) chunk\s+(-?\d+)\s+rc (-?\d+)\s+parity (\S+).*?"
                  r"mean ([\d.]+) ms.*?([\d.]+) TFLOP/s")


def decode(var, bal, chunk):
    if 100 <= var < 150:
        return {"algorithm": "dmma", "warps_per_cta": WPC[(var - 100) // 10], "flush": var % 10, "chunk": chunk, "align_runs": bool(bal & 2)}
    if var in (200, 201):
        return {"algorithm": "tiny", "warps_per_cta": 8 if var == 200 else 4, "flush": 0, "chunk": chunk, "align_runs": False}
    return None  # variant 9 and experiments: not a database configuration


def main():
    src = sys.argv[1]
    source = "autotuned: " + os.path.relpath(os.path.abspath(src), ROOT)
    if "--source" in sys.argv:
        source = sys.argv[sys.argv.index("--source") + 1]
    best, base, bad = {}, {}, 0
    for line in open(src):
        m = LINE.match(line)
        if not m:
            continue
        mm, nn, kk, spec, var, bal, chunk, rc, parity, ms, tf = m.groups()
        key = (int(mm), int(nn), int(kk))
        var, bal, chunk, tf = int(var), int(bal), int(chunk), float(tf)
        if parity not in ("exact", "reference") or int(rc) != 0:
            bad += 1
            continue
        if var == 9:
            base[key] = tf
        cfg = decode(var, bal, chunk)
        if cfg is None:
            continue
        if key not in best or tf > best[key][0]:
            best[key] = (tf, cfg)
    db_path = os.path.join(ROOT, "dbcsr_b200", "parameters", "parameters_B200.json")
    old = {(r["m"], r["n"], r["k"]): r for r in json.load(open(db_path))}
    out = []
    for key in sorted(set(old) | set(best)):
        if key in best:
            tf, cfg = best[key]
            rec = {"m": key[0], "n": key[1], "k": key[2], "algorithm": cfg["algorithm"], "warps_per_cta": cfg["warps_per_cta"], "stages": 0,
                   "flush": cfg["flush"], "chunk": cfg["chunk"], "align_runs": cfg["align_runs"], "perf": round(tf * 1e3, 1),
                   "perf_default": round(base.get(key, 0.0) * 1e3, 1), "source": source}
        else:
            rec = old[key]
        out.append(rec)
    print("%d triplets tuned, %d kept from the old database, %d runs rejected (parity / rc)" % (len(best), len(out) - len(best), bad))
    for key in sorted(best):
        tf, cfg = best[key]
        print("  %2d %2d %2d  %7.2f TFLOP/s (default %6.2f)  %s wpc %d flush %d chunk %d %s" % (key + (tf, base.get(key, 0.0), cfg["algorithm"], cfg["warps_per_cta"],
                                                                                            cfg["flush"], cfg["chunk"], "aligned" if cfg["align_runs"] else "")))
    if "--dry" not in sys.argv:
        with open(db_path, "w") as f:
            f.write("[\n" + ",\n".join(json.dumps(r) for r in out) + "\n]\n")


if __name__ == "__main__":
    main()
