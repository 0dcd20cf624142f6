#!/usr/bin/env python
"""Ablation of the split-fp16 MMA products per convolution (VERDICT r1 "next" 3).

For a synthetic population of one BASELINE workload the exact-fp32 SIMT path gives the ground truth (it is within 5e-6
of the reference's own fitness, profiles/r1/parity_vs_reference_l.txt).  Every configuration below is one setting of
`eig_set_option` ("passes.*" masks: bit 0 a_lo*w_hi, bit 1 a_hi*w_lo, bit 2 a_hi*w_hi; "early_until"/"early_mask": cheaper
products on the first PredNet steps only).  The FLOOR row is the same exact-fp32 kernel with the nine taps summed in
reverse order: what any other, equally exact fp32 implementation (cuDNN vs MKL vs this one) does to the frames and to the
fitness.  Reported per configuration: device ms per evaluation chunk, fraction of frame bytes that differ from the fp32
path (and the largest difference), and the fitness error against the fp32 path (relative; `north_star` tolerance 1e-3).

  python profiles/experiments/pass_ablation.py --workload c3 --total 512 --chunk 64 --set mixes
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from evolutionary_illusion_generator_b200 import _lib, engine as E, weights as W  # noqa: E402

UPPER = [("passes.L2", 4), ("passes.L3", 4), ("passes.A2", 4), ("passes.A3", 4), ("passes.P2", 4), ("passes.P3", 4)]
NAMES = {6: "drop a_lo*w_hi", 5: "drop a_hi*w_lo", 4: "hi*hi only"}


def config_sets(which):
    sets = {}
    sets["mixes"] = [
        ("3-pass everywhere (precision 0)", []),
        ("all convs: drop a_lo*w_hi", [("passes.all", 6)]),
        ("all convs: drop a_hi*w_lo", [("passes.all", 5)]),
        ("all convs: hi*hi only (precision 2)", [("passes.all", 4)]),
        ("layers 2+3 hi*hi only, layer 1 3-pass (precision 1)", UPPER),
        ("layers 2+3 hi*hi only + L1 drop a_lo*w_hi", UPPER + [("passes.L1", 6)]),
        ("layers 2+3 hi*hi only + L1 drop a_hi*w_lo", UPPER + [("passes.L1", 5)]),
        ("layers 2+3 hi*hi only + L1 hi*hi only", UPPER + [("passes.L1", 4)]),
        ("layers 2+3 hi*hi only + A1 hi*hi only", UPPER + [("passes.A1", 4)]),
        ("layers 2+3 hi*hi only + P1/Z hi*hi only", UPPER + [("passes.P1", 4)]),
        ("layers 2+3 hi*hi only + L1, A1 hi*hi only", UPPER + [("passes.L1", 4), ("passes.A1", 4)]),
    ]
    L3 = [("passes.L3", 4), ("passes.A3", 4), ("passes.P3", 4)]
    sets["layer3"] = [
        ("3-pass everywhere (precision 0)", []),
        ("layer 3 (L3, A3, P3) hi*hi only", L3),
        ("layer 3 + P2 hi*hi only", L3 + [("passes.P2", 4)]),
        ("layer 3 + P2 hi*hi only, L2 drop a_lo*w_hi", L3 + [("passes.P2", 4), ("passes.L2", 6)]),
        ("layer 3 + P2 hi*hi only, L2 drop a_hi*w_lo", L3 + [("passes.P2", 4), ("passes.L2", 5)]),
        ("layer 3 + P2 + A2 hi*hi only", L3 + [("passes.P2", 4), ("passes.A2", 4)]),
        ("layer 3 + P2 + A2 hi*hi only, L2 drop a_lo*w_hi", L3 + [("passes.P2", 4), ("passes.A2", 4), ("passes.L2", 6)]),
        ("layers 2+3 hi*hi only, layer 1 3-pass", UPPER),
        ("3-pass everywhere, repeated (run-to-run: identical bits expected)", []),
    ]
    single = [("3-pass everywhere (precision 0)", [])]
    for tgt in ("L1", "L2", "L3", "A1", "A2", "A3", "P1", "P2", "P3"):
        for m in (6, 5, 4):
            single.append(("%s: %s" % (tgt, NAMES[m]), [("passes." + tgt, m)]))
    sets["single"] = single
    early = [("3-pass everywhere (precision 0)", [])]
    for T in (5, 10, 14, 17, 19):
        for m in (4, 6, 5):
            early.append(("steps < %d: %s" % (T, NAMES[m]), [("early_until", T), ("early_mask", m)]))
    sets["early"] = early
    out = []
    for w in which.split(","):
        out += [c for c in sets[w] if c[0] not in [o[0] for o in out]]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--total", type=int, default=256, help="genomes in the sample")
    ap.add_argument("--chunk", type=int, default=64, help="genomes per evaluation")
    ap.add_argument("--set", default="mixes", help="comma list of: mixes, layer3, single, early")
    ap.add_argument("--start", type=int, default=0, help="index of the first synthetic genome")
    args = ap.parse_args()
    preset, c_dim, ch, w, h, structure, _, _ = bench.WORKLOADS[args.workload]
    chunk, total = args.chunk, args.total
    eng = E.Engine(w, h, ch, chunk, device=0)
    eng.set_grid(structure)
    eng.load_weights(W.synthetic_predictor_weights(w, h, ch, seed=0))
    residents = []
    for c0 in range(0, total, chunk):
        _, _, progs = bench.build_population(preset, c_dim, min(chunk, total - c0), args.start + c0)
        residents.append(eng.upload_programs(progs))

    def run():
        fits, frames, ms = [], [], []
        for res in residents:
            n = res[2]
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eng.evaluate_resident(res, structure)            # warm (graph capture happens on the second use of a key)
            a.record()
            fit = eng.evaluate_resident(res, structure)
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
            fits.append(fit.cpu().numpy().copy())
            frames.append(eng.debug_buffers(n)["frames"][:2].copy())
        return np.concatenate(fits), frames, float(np.median(ms))

    def reset():
        eng.set_option("precision", 0)
        eng.set_option("early_until", 0)
        eng.set_option("early_mask", 7)
        eng.set_option("simt_reverse_taps", 0)

    eng.set_conv_mode(_lib.CONV_SIMT)
    reset()
    f_ref, fr_ref, ms_simt = run()

    print("# MMA-product ablation, workload %s: %d synthetic genomes (index %d..), %d per evaluation (%dx%d, channels %s), "
          "synthetic predictor weights seed 0\n" % (args.workload, total, args.start, chunk, w, h, list(ch)))
    print("Ground truth: exact-fp32 SIMT path of the same library (%.1f ms per evaluation of %d genomes).  `ms` = median device "
          "time of one resident evaluation of %d genomes (CUDA events, graph replay).  Frame columns: the two frames handed "
          "to the flow stage, %d bytes in total.  Outliers = genomes whose fitness differs from the fp32 path by more than 1e-3 "
          "relative (ids listed).\n" % (ms_simt, chunk, chunk, sum(f.size for f in fr_ref)))
    print("| configuration | ms | frame bytes differing | max LSB | fitness rel. err p50 | p90 | p99 | max | > 1e-3 | > 1e-2 | outlier ids |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")

    def report(name, f, fr, ms):
        nd = sum(int((a.astype(np.int16) != b.astype(np.int16)).sum()) for a, b in zip(fr, fr_ref))
        mx = max(int(np.abs(a.astype(np.int16) - b.astype(np.int16)).max()) for a, b in zip(fr, fr_ref))
        nb = sum(a.size for a in fr_ref)
        rel = np.abs(f - f_ref) / np.maximum(np.abs(f_ref), 1e-12)
        rel[(f == 0) & (f_ref == 0)] = 0.0
        rel = np.where(np.isnan(f) & np.isnan(f_ref), 0.0, rel)
        rel = np.where(np.isnan(rel), np.inf, rel)
        out = np.nonzero(rel > 1e-3)[0]
        print("| %s | %.2f | %.2e | %d | %.1e | %.1e | %.1e | %.1e | %d / %d | %d | %s |" % (
            name, ms, nd / nb, mx, np.quantile(rel, 0.5), np.quantile(rel, 0.9), np.quantile(rel, 0.99), rel.max(),
            len(out), total, int((rel > 1e-2).sum()), " ".join(str(int(i)) for i in out[:16])))
        sys.stdout.flush()

    eng.set_option("simt_reverse_taps", 1)
    f, fr, ms = run()
    report("FLOOR: exact fp32, taps summed in reverse order", f, fr, ms)
    eng.set_option("simt_reverse_taps", 0)
    eng.set_conv_mode(_lib.CONV_TC)
    for name, opts in config_sets(args.set):
        reset()
        for k, v in opts:
            eng.set_option(k, v)
        f, fr, ms = run()
        report(name, f, fr, ms)
    reset()
    eng.close()


if __name__ == "__main__":
    main()
