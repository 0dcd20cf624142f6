import sys, time
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import torch
from evolutionary_illusion_generator_b200 import _lib, engine as E, genome as G, weights as W
import bench
w, h, ch = 160, 120, (1, 16, 32, 64)
wts = W.synthetic_predictor_weights(w, h, ch, seed=0)
def make(n, start):
    eng = E.Engine(w, h, ch, n, device=0); eng.set_conv_mode(_lib.CONV_TC); eng.set_grid(1); eng.load_weights(wts)
    _, _, progs = bench.build_population("circles_bw", 1, n, start)
    return eng, eng.upload_programs(progs), torch.empty((n,), dtype=torch.float64, device='cuda')
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / reps * 1e3
e32, r32, f32 = make(32, 0)
ms1 = timeit(lambda: e32.evaluate_resident(r32, 1, out=f32))
print("1 x 32 genomes: %.2f ms" % ms1)
for parts in (2, 4):
    n = 32 // parts
    engs = [make(n, i * n) for i in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    def run():
        for (e, r, f), s in zip(engs, streams):
            with torch.cuda.stream(s):
                e.evaluate_resident(r, 1, out=f)
    ms = timeit(run)
    got = torch.cat([f for _, _, f in engs]).cpu().numpy()
    print("%d x %d genomes on %d streams: %.2f ms, same fitness: %s" % (parts, n, parts, ms, bool((got == f32.cpu().numpy()).all())))
