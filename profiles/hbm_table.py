"""Achieved HBM GB/s of the memory-side kernels of workload C2 (32 gray genomes, 160x120, channels 1,16,32,64):
algorithmic bytes per launch (what the kernel must read + write once, from the buffer shapes in DESIGN.md section 3)
divided by the kernel's average duration in the ncu launch list (`gpu__time_duration.sum`, profiles/r1/launches_c2_*.txt).
This is a derived table, not a `dram__bytes` capture: ncu --set full was spent on the convolution kernel this round.
Usage: python profiles/hbm_table.py profiles/r1/launches_c2_m_final.txt"""
import json
import os
import re
import sys

B, H, W, C0, C1 = 32, 120, 160, 1, 16
px0, px1 = B * H * W, B * (H // 2) * (W // 2)
f32 = 4
BYTES = {   # kernel name prefix -> (algorithmic bytes per launch, what is counted)
    "eig::cppn_render_kernel": (px0 * C0 * (1 + f32) + 2 * H * W * 8, "u8 image + fp32 input out, two fp64 grid planes in (fp64-ALU bound)"),
    "eig::l0_conva1_kernel": (2 * px0 * C0 * f32 + px1 * C1 * f32 + px1 * 2 * C1 * f32, "x, P0, P1 in; split-fp16 E1 (2*C1 ch x 4 B) out"),
    "eig::l0_lstm_kernel": (2 * px0 * C0 * f32 + px1 * 16 * C0 * f32 + 2 * px0 * C0 * f32 + 2 * px0 * C0 * f32, "x, P0, Z, h0, c0 in; h0, c0 out"),
    "eig::l0_convp_kernel": (2 * px0 * C0 * f32, "h0 in, P0 out"),
    "eig::quantize_gray_kernel": (px0 * C0 * f32 + 2 * px0, "P0 in, u8 frame + u8 gray out"),
    "eig::reset_state_kernel": (None, "15 state regions cleared"),
    "eig::min_eig_kernel": (px0 * (1 + f32), "u8 gray in, fp32 eigenvalue map out"),
    "eig::scharr_kernel": (px0 * (1 + 4), "u8 level in, 2 x int16 derivatives out (level 0)"),
    "eig::pyr_down_kernel": (2 * px0 + 2 * px0 // 4, "two u8 level-0 images in, two level-1 images out"),
}
# reset: X1..X3 even buffers (both fp16 planes = 4 B/elem), h0, c_n, P_n
ctot = {1: 3 * 16 + 32, 2: 3 * 32 + 64, 3: 3 * 64}
ch = {0: 1, 1: 16, 2: 32, 3: 64}
reset = sum(B * (H >> n) * (W >> n) * ctot[n] * 4 for n in (1, 2, 3)) + px0 * f32 + sum(2 * B * (H >> n) * (W >> n) * ch[n] * f32 for n in range(4))
BYTES["eig::reset_state_kernel"] = (reset, BYTES["eig::reset_state_kernel"][1])

peaks = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))
peak = peaks.get("hbm_gbs", 6550.0)
print("| kernel | launch shape | avg duration (ncu) | algorithmic bytes | achieved GB/s | of %.0f GB/s | counted |" % peak)
print("|---|---|---|---|---|---|---|")
seen = set()
for line in open(sys.argv[1]):
    m = re.match(r"(\S+)\s+\((\d+), \d+, \d+\)\s+\((\d+), (\d+), \d+\)\s+n=\s*\d+\s+[\d.]+ us\s+[\d.]+% avg\s+([\d.]+) us", line)
    if not m:
        continue
    name = m.group(1).split("<")[0]
    if name not in BYTES or name in seen:
        continue
    seen.add(name)
    nbytes, what = BYTES[name]
    us = float(m.group(5))
    gbs = nbytes / us / 1e3
    print("| `%s` | %s x (%s,%s) | %.1f us | %.2f MB | %.0f | %.2f | %s |" % (m.group(1), m.group(2), m.group(3), m.group(4), us, nbytes / 1e6, gbs, gbs / peak, what))
