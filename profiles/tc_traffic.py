#!/usr/bin/env python
"""profiles/r2/tc_traffic.json (read by bench.py for `roofline.traffic`): mean measured DRAM bytes per launch of the
dominant kernel (conv3x3_tc_kernel) over ALL its launches of one evaluation, from the ncu metric logs
(`--metrics dram__bytes_read.sum,dram__bytes_write.sum,...`, profiles/kernel_metrics.py format).

  python profiles/tc_traffic.py profiles/r2/tc_traffic.json c2=gpurun_out/x/kernel_metrics_c2.csv.gz c3=...
"""
import collections
import csv
import gzip
import json
import sys


def load(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rd:
        d = per.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]]})
        v = float(r[ix["Metric Value"]].replace(",", ""))
        sc = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ix["Metric Unit"]], 1.0)
        d[r[ix["Metric Name"]]] = v * sc
    return [d for d in per.values() if "conv3x3_tc_kernel" in d["name"]]


out = {}
for arg in sys.argv[2:]:
    name, path = arg.split("=")
    L = load(path)
    tot = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in L)
    us = sum(d["gpu__time_duration.sum"] for d in L)
    out[name] = {"launches": len(L), "mean_dram_bytes_per_launch": tot / len(L), "dram_bytes_per_evaluation": tot,
                 "kernel_us_per_evaluation": us,
                 "mean_tensor_pipe_active_pct": sum(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] * d["gpu__time_duration.sum"] for d in L) / us,
                 "source": path, "note": "conv3x3_tc_kernel, every launch of one whole evaluation (21 PredNet steps), ncu --clock-control none"}
json.dump(out, open(sys.argv[1], "w"), indent=1)
print({k: (round(v["mean_dram_bytes_per_launch"] / 1e6, 1), v["launches"], round(v["mean_tensor_pipe_active_pct"], 1)) for k, v in out.items()})
