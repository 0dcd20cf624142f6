"""Whole path in exact-fp32 (SIMT) mode on 4 genomes, for `compute-sanitizer --tool racecheck` (the tool does not model the
asynchronous-proxy writes of TMA / tcgen05, so the tensor-core kernel is checked with memcheck and the parity tests instead)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from evolutionary_illusion_generator_b200 import _lib, engine as E, genome as G, weights as W

w, h, ch = 64, 64, (1, 16, 32, 64)
eng = E.Engine(w, h, ch, 4)
eng.set_conv_mode(_lib.CONV_SIMT)
eng.set_grid(1)
eng.load_weights(W.synthetic_predictor_weights(w, h, ch, seed=0))
cfg = G.make_config(2, 1)
progs = [G.flatten_genome(G.synthetic_genome("circles_bw", i), cfg, n_outputs=1) for i in range(4)]
print("fitness", eng.evaluate(progs, 1))
