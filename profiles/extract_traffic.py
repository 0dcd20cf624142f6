#!/usr/bin/env python
"""Per-launch DRAM traffic of the dominant kernel from `ncu -i X.ncu-rep --page raw --csv` dumps.
usage: extract_traffic.py out.json name=raw.csv [name=raw.csv ...]"""
import csv, json, sys
out = {}
for arg in sys.argv[2:]:
    name, path = arg.split("=")
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    ip = hdr.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
    tscale = {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}
    L = []
    for r in rows[2:]:
        L.append(dict(dram_bytes=float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]],
                      us=float(r[it]) * tscale[units[it]], tensor_pipe_active_pct=float(r[ip])))
    out[name] = dict(launches=L, mean_dram_bytes_per_launch=sum(x["dram_bytes"] for x in L) / len(L),
                     note="conv3x3_tc_kernel, the 8 launches of one PredNet step (ConvA2 ConvA3 LSTM3 LSTM2 LSTM1 ConvP1+Z ConvP2 ConvP3), ncu --set full --clock-control none")
json.dump(out, open(sys.argv[1], "w"), indent=1)
print({k: round(v["mean_dram_bytes_per_launch"] / 1e6, 2) for k, v in out.items()}, "MB per launch")
