#!/usr/bin/env python
"""Per-kernel table from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active...,sm__throughput... --csv` log (long format: one row per launch and metric).

  python profiles/kernel_metrics.py gpurun_out/kernel_metrics_c2.csv.gz "title" > profiles/r2/kernel_metrics_c2.md

Launches are grouped by (kernel name, grid, block); per group: launches, mean duration, mean measured DRAM bytes
(read + write), achieved DRAM GB/s = bytes / duration, its fraction of the measured HBM peak (MEASURED_PEAKS.json),
tensor-pipe active % and SM throughput %.  ncu serialises the launches and starts each one with cold L1 (L2 keeps what
the previous kernels left), so the durations are per-kernel figures, not shares of an overlapped step."""
import collections
import csv
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path, title = sys.argv[1], sys.argv[2]
    op = gzip.open if path.endswith(".gz") else open
    rows = []
    with op(path, "rt") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rd:
        key = r[ix["ID"]]
        d = per.setdefault(key, {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]]})
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "%": 1.0}.get(unit, 1.0)
        d[r[ix["Metric Name"]]] = v * scale
    groups = collections.OrderedDict()
    for d in per.values():
        name = d["name"].split("(")[0].replace("eig::", "")
        groups.setdefault((name, d["grid"], d["block"]), []).append(d)
    peak = 6543.7
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    print("# %s\n" % title)
    print("`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,"
          "sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none`; DRAM bytes are MEASURED (`dram__bytes_*`), "
          "per launch; GB/s = bytes / duration; peak = %.1f GB/s (MEASURED_PEAKS.json).  Made by `profiles/kernel_metrics.py`.\n" % peak)
    print("| kernel | grid x block | launches | mean us | total ms | DRAM read MB | DRAM write MB | achieved GB/s | of HBM peak | tensor pipe % | SM throughput % |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    tot = sum(d.get("gpu__time_duration.sum", 0.0) for d in per.values())
    for (name, grid, block), L in sorted(groups.items(), key=lambda kv: -sum(d.get("gpu__time_duration.sum", 0.0) for d in kv[1])):
        n = len(L)
        us = sum(d.get("gpu__time_duration.sum", 0.0) for d in L) / n
        rd_b = sum(d.get("dram__bytes_read.sum", 0.0) for d in L) / n
        wr_b = sum(d.get("dram__bytes_write.sum", 0.0) for d in L) / n
        tp = sum(d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) for d in L) / n
        smt = sum(d.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", 0.0) for d in L) / n
        gbs = (rd_b + wr_b) / us / 1e3 if us > 0 else 0.0
        print("| `%s` | %s x %s | %d | %.1f | %.2f | %.2f | %.2f | %.0f | %.3f | %.1f | %.1f |" % (
            name, grid, block, n, us, n * us / 1e3, rd_b / 1e6, wr_b / 1e6, gbs, gbs / peak, tp, smt))
    print("\nAll launches: %d, %.2f ms in total." % (len(per), tot / 1e3))


if __name__ == "__main__":
    main()
