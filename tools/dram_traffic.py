#!/usr/bin/env python3
"""DRAM traffic of ONE multiply of BASELINE config 2 through the FP64 stack kernel, for DBCSR's own stacks and for the device
builder's tile-ordered stacks (cfg.dev_tile).  Run under
  ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      --csv --log-file gpurun_out/dram_traffic.csv python tools/dram_traffic.py
(only the two drains between cudaProfilerStart/Stop are profiled); tools/dram_traffic.py --digest <csv> sums the launches."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def digest(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    i_name, i_metric, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    i_id = hdr.index("ID")
    per = {}
    for r in rows[1:]:
        if "smm_dmma_kernel" not in r[i_name]:
            continue
        v = float(r[i_val].replace(",", ""))
        u = r[i_unit]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(u, 1.0)
        per.setdefault(int(r[i_id]), {})[r[i_metric]] = v * scale
    ids = sorted(per)
    half = len(ids) // 2
    out = {}
    for name, sel in (("dbcsr_order", ids[:half]), ("tile_order", ids[half:])):
        rd = sum(per[i].get("dram__bytes_read.sum", 0.0) for i in sel)
        wr = sum(per[i].get("dram__bytes_write.sum", 0.0) for i in sel)
        tm = sum(per[i].get("gpu__time_duration.sum", 0.0) for i in sel)
        out[name] = {"launches": len(sel), "dram_read_GB": rd * 1e-9, "dram_write_GB": wr * 1e-9, "dram_total_GB": (rd + wr) * 1e-9,
                     "sum_of_isolated_launch_times_ms": tm * 1e3}
    out["note"] = ("one multiply of cfg2 (1000x1000 grid, 23x23 blocks, 10 %): 334 launches each; ncu serialises the launches (cold-cache, no PDL overlap), "
                   "so the byte counts are the evidence, not the times; compulsory traffic = A + B once + C written once = 9.3 GB incl. the read of the zeroed C lines")
    return out


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--digest":
        print(json.dumps(digest(sys.argv[2]), indent=1))
        return
    import torch

    import bench
    from dbcsr_b200 import lib as acclib

    acc = acclib.Acc(0)
    s = acc.stream_create("traffic", 0)
    torch.cuda.init()
    runs = [bench.Fp64Run(acc, "cfg2", None, s), bench.Fp64Run(acc, "cfg2", None, s, dev_tile=64)]
    for r in runs:  # warm the kernels up outside the profiled range
        acc.memset_zero(r.d_cs[0], s)
        r.drain(r.d_cs[0])
    acc.stream_sync(s)
    torch.cuda.profiler.start()
    for r in runs:
        acc.memset_zero(r.d_cs[0], s)
        r.drain(r.d_cs[0])
        acc.stream_sync(s)
    torch.cuda.profiler.stop()
    for r in runs:
        r.close()
    print("ok", [len(r.stacks) for r in runs])


if __name__ == "__main__":
    main()
