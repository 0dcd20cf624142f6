"""Wall-clock of one generation seen from the genome objects (SURVEY.md §8 f row 4): flatten up front + eig_eval_host
against Engine.evaluate_streamed (flatten of chunk k+1 overlapped with the GPU evaluation of chunk k), with a cold and a
warm ProgramCache.  Usage: python profiles/experiments/host_pipeline_exp.py [c2|c3]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
import torch
from evolutionary_illusion_generator_b200 import _lib, engine as E, genome as G, weights as W

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
preset, c, ch, n = {"c2": ("circles_bw", 1, (1, 16, 32, 64), 32), "c3": ("circles", 3, (3, 48, 96, 192), 128)}[which]
w, h = 160, 120
eng = E.Engine(w, h, ch, n, device=0)
eng.set_conv_mode(_lib.CONV_TC)
eng.set_grid(1)
eng.load_weights(W.synthetic_predictor_weights(w, h, ch, seed=0))
cfg = G.make_config(2, G.NEAT_PRESETS[preset]["num_outputs"])
pop = G.synthetic_population(preset, n, evolved=True)


def upfront():
    progs = [G.flatten_genome(g, cfg, n_outputs=c) for _, g in pop]
    return eng.evaluate(progs, 1)


def streamed_cold():
    return eng.evaluate_streamed(pop, lambda gid, g: G.flatten_genome(g, cfg, n_outputs=c), 1).cpu().numpy()


cache = G.ProgramCache()


def streamed_warm():
    return eng.evaluate_streamed(pop, lambda gid, g: cache.get(gid, g, cfg, c), 1).cpu().numpy()


def flatten_only():
    return [G.flatten_genome(g, cfg, n_outputs=c) for _, g in pop]


def timeit(fn, reps):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        out = fn()
        ts.append((time.perf_counter() - t) * 1e3)
    return float(np.median(ts)), out


reps = 20 if which == "c2" else 8
t_flat, _ = timeit(flatten_only, reps)
t_a, fa = timeit(upfront, reps)
t_b, fb = timeit(streamed_cold, reps)
t_c, fc = timeit(streamed_warm, reps)
print("%s: %d genomes, chunk %d" % (which, n, eng.stream_chunk(n)))
print("  flatten only                               %8.2f ms" % t_flat)
print("  flatten up front + eig_eval_host           %8.2f ms  (%.0f evals/s)" % (t_a, n / t_a * 1e3))
print("  evaluate_streamed, every genome flattened  %8.2f ms  (%.0f evals/s)" % (t_b, n / t_b * 1e3))
print("  evaluate_streamed, program cache warm      %8.2f ms  (%.0f evals/s)" % (t_c, n / t_c * 1e3))
print("  same fitness bits: %s" % bool(np.array_equal(fa, fb, equal_nan=True) and np.array_equal(fa, fc, equal_nan=True)))
for chunk in sorted({n, n // 2, n // 4, n // 8}, reverse=True):
    t_cold, f1 = timeit(lambda: eng.evaluate_streamed(pop, lambda gid, g: G.flatten_genome(g, cfg, n_outputs=c), 1, chunk=chunk).cpu().numpy(), reps)
    t_warm, f2 = timeit(lambda: eng.evaluate_streamed(pop, lambda gid, g: cache.get(gid, g, cfg, c), 1, chunk=chunk).cpu().numpy(), reps)
    print("  chunk %3d: every genome flattened %8.2f ms, cache warm %8.2f ms, same bits %s"
          % (chunk, t_cold, t_warm, bool(np.array_equal(fa, f1, equal_nan=True) and np.array_equal(fa, f2, equal_nan=True))))
