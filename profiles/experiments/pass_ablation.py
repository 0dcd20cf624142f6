#!/usr/bin/env python
"""Ablation of the split-fp16 MMA products per convolution (VERDICT r1 "next" 3).

For a synthetic population of one BASELINE workload the exact-fp32 SIMT path gives the ground truth (it is within 5e-6
of the reference's own fitness, profiles/r1/parity_vs_reference_l.txt).  Every configuration below is one setting of
`eig_set_option` ("passes.*" masks: bit 0 a_lo*w_hi, bit 1 a_hi*w_lo, bit 2 a_hi*w_hi; "early_until"/"early_mask": cheaper
products on the first PredNet steps only).  Reported per configuration: device ms per evaluation, fraction of frame
bytes that differ from the fp32 path (and the largest difference), and the fitness error against the fp32 path
(relative, the `north_star` tolerance is 1e-3).

  python profiles/experiments/pass_ablation.py --workload c3 --pop 64 > profiles/r2/pass_ablation_c3.md
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--pop", type=int, default=64)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--configs", default="all")
    args = ap.parse_args()
    preset, c_dim, ch, w, h, structure, _, _ = bench.WORKLOADS[args.workload]
    pop = args.pop
    eng = E.Engine(w, h, ch, pop, device=0)
    eng.set_grid(structure)
    eng.load_weights(W.synthetic_predictor_weights(w, h, ch, seed=0))
    _, _, progs = bench.build_population(preset, c_dim, pop, 0)
    resident = eng.upload_programs(progs)

    def run(reps):
        fit = None
        ms = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fit = eng.evaluate_resident(resident, structure)
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        bufs = eng.debug_buffers(pop)
        return fit.cpu().numpy().copy(), bufs["frames"][:2].copy(), min(ms)

    eng.set_conv_mode(_lib.CONV_SIMT)
    f_ref, fr_ref, ms_simt = run(1)
    eng.set_conv_mode(_lib.CONV_TC)

    def reset():
        eng.set_option("passes.all", 7)
        eng.set_option("early_until", 0)
        eng.set_option("early_mask", 7)

    configs = [("3-pass everywhere (shipped)", [])]
    for m, nm in ((6, "drop a_lo*w_hi"), (5, "drop a_hi*w_lo"), (4, "hi*hi only")):
        configs.append(("all convs: %s" % nm, [("passes.all", m)]))
    for tgt in ("L1", "L2", "L3", "L", "A", "P", "A2", "A3", "P1", "P2", "P3"):
        for m, nm in ((6, "drop a_lo*w_hi"), (5, "drop a_hi*w_lo"), (4, "hi*hi only")):
            configs.append(("%s: %s" % (tgt, nm), [("passes." + tgt, m)]))
    for T in (5, 10, 14, 17, 19):
        for m, nm in ((4, "hi*hi only"), (6, "drop a_lo*w_hi"), (5, "drop a_hi*w_lo")):
            configs.append(("steps < %d: %s" % (T, nm), [("early_until", T), ("early_mask", m)]))
    if args.configs != "all":
        keep = args.configs.split(",")
        configs = [c for c in configs if any(k in c[0] for k in keep)]

    print("# MMA-product ablation, workload %s, %d genomes (%dx%d, channels %s), synthetic predictor weights seed 0\n"
          % (args.workload, pop, w, h, list(ch)))
    print("Ground truth: exact-fp32 SIMT path of the same library (%.2f ms per evaluation).  `ms` = best of %d resident "
          "evaluations (CUDA events, graphs on).  Frame columns: the two frames handed to the flow stage.\n"
          % (ms_simt, args.reps))
    print("| configuration | ms | frame bytes differing | max LSB | fitness rel. err max | median | genomes > 1e-3 | genomes > 1e-2 |")
    print("|---|---|---|---|---|---|---|---|")
    for name, opts in configs:
        reset()
        for k, v in opts:
            eng.set_option(k, v)
        run(2)   # first use direct, second captures the graph
        f, fr, ms = run(args.reps)
        d = fr.astype(np.int32) - fr_ref.astype(np.int32)
        denom = np.maximum(np.abs(f_ref), 1e-12)
        rel = np.abs(f - f_ref) / denom
        rel[(f == 0) & (f_ref == 0)] = 0.0
        rel = np.where(np.isnan(f) & np.isnan(f_ref), 0.0, rel)
        print("| %s | %.2f | %.2e | %d | %.2e | %.2e | %d / %d | %d |" % (
            name, ms, float((d != 0).mean()), int(np.abs(d).max()), float(np.nanmax(rel)), float(np.nanmedian(rel)),
            int((rel > 1e-3).sum()), pop, int((rel > 1e-2).sum())))
        sys.stdout.flush()
    reset()
    eng.close()


if __name__ == "__main__":
    main()
