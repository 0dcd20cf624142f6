#!/usr/bin/env python
"""A handful of whole-path evaluations of one BASELINE workload - the command ncu wraps for the per-kernel metric lists
(profiles/r2/kernel_metrics_*.md).  EIG_NO_GRAPH=1 makes every kernel a plain launch."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from evolutionary_illusion_generator_b200 import _lib, engine as E, weights as W  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--pop", type=int, default=0)
ap.add_argument("--evals", type=int, default=2)
ap.add_argument("--conv", default="tc")
args = ap.parse_args()
preset, c_dim, ch, w, h, structure, pop, _ = bench.WORKLOADS[args.workload]
pop = args.pop or pop
eng = E.Engine(w, h, ch, pop, device=0)
eng.set_conv_mode(_lib.CONV_TC if args.conv == "tc" else _lib.CONV_SIMT)
eng.set_grid(structure)
eng.load_weights(W.synthetic_predictor_weights(w, h, ch, seed=0))
_, _, progs = bench.build_population(preset, c_dim, pop, 0)
resident = eng.upload_programs(progs)
for _ in range(args.evals):
    fit = eng.evaluate_resident(resident, structure)
torch.cuda.synchronize()
print("fitness checksum", float(fit.nansum()), "launches", eng.lib.eig_launch_count())
eng.close()
