#!/usr/bin/env python3
"""SASS census of the shipped library -> profiles/sass_r02.txt (run after a build; needs cuobjdump and c++filt).
Usage: python tools/sass_census.py [lib.so] [out.txt]"""
import collections, re, subprocess, sys

lib = sys.argv[1] if len(sys.argv) > 1 else "dbcsr_b200/lib/libdbcsr_acc_b200.so"
out_path = sys.argv[2] if len(sys.argv) > 2 else "profiles/sass_r02.txt"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
pat = re.compile(r"\b(REDG|ATOMG|UBLKRED|UBLKCP|UTCHMMA|UTCCP|LDTM|DMMA|UTCBAR|SYNCS|UTMALDG|UTMASTG|DFMA|LDS|STS|LDG|STG|ELECT|UTCATOMSWS|ACQBULK|DADD|DMUL)[A-Za-z0-9_.]*")
plain = ("LDS", "STS", "LDG", "STG", "DADD", "DMUL")
out = ["# SASS census of %s (cuobjdump -sass, sm_100a); regenerate with tools/sass_census.py" % lib]
tot = collections.Counter(m.group(0) for m in pat.finditer(txt))
out.append("\n## whole library (%d kernels)" % (len(funcs) - 1))
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    if k.split(".")[0] not in plain:
        out.append("%8d  %s" % (v, k))
out.append("\n# proof points: DMMA.8x8x4 = FP64 tensor pipe (mma.sync.m8n8k4.f64); UBLKCP.S.G = cp.async.bulk global->shared (1-D TMA);")
out.append("# UBLKRED.G.S.ADD.F64 = cp.reduce.async.bulk shared->global add.f64 (bulk flush); UTCHMMA = tcgen05.mma; UTCCP = tcgen05.cp;")
out.append("# LDTM = tcgen05.ld; UTCBAR = tcgen05.commit; SYNCS.* = mbarrier; no UTMALDG/UTMASTG: no tensor-map TMA is used (blocks are 1-D runs)")
out.append("\n## per kernel (demangled name : selected mnemonics)")
names = []
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    c = collections.Counter(
        m.group(0).split(".")[0] if m.group(0).split(".")[0] in plain + ("SYNCS",) else m.group(0) for m in pat.finditer(f))
    names.append((name, c))
dem = subprocess.run(["c++filt"], input="\n".join(n for n, _ in names), capture_output=True, text=True).stdout.split("\n")
for (n, c), d in zip(names, dem):
    d = re.sub(r"\(.*", "", d)
    out.append("%s : %s" % (d or n, ", ".join("%s=%d" % (k, v) for k, v in sorted(c.items()))))
open(out_path, "w").write("\n".join(out) + "\n")
print("wrote", out_path, len(out), "lines")
