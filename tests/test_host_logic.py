"""Host-side logic and the C-ABI surface (no compute calls): CPU only."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from evolutionary_illusion_generator_b200 import _lib, genome as G, grid as PG, runtime, weights as W
from flat_interp import render_flat
from oracle import cppn as OC, grid as OG


def test_grid_matches_oracle_all_structures():
    for s in range(4):
        for (w, h) in [(160, 120), (64, 64), (320, 240)]:
            if s == 0 and w % 10:
                continue
            a, b = OG.create_grid(s, w, h, 10), PG.create_grid(s, w, h, 10)
            assert np.array_equal(a["x_mat"], b["x_mat"]) and np.array_equal(a["y_mat"], b["y_mat"])


def test_enhanced_grid_matches_oracle_and_reference_digest():
    """`enhanced_image_grid` (generate_illusion.py:121-193): vectorised product version == scalar oracle
    on small mosaics of every structure, and == the reference at 800x800 through the digest recorded by
    tests/golden/make_golden.py from the reference's own function."""
    import hashlib
    import json
    from conftest import GOLDEN
    for st, w, h in ((1, 96, 96), (3, 90, 120), (0, 60, 60), (2, 75, 75)):
        a, b = PG.enhanced_image_grid(w, h, st), OG.enhanced_image_grid(w, h, st)
        assert np.array_equal(a["x_mat"], b["x_mat"]) and np.array_equal(a["y_mat"], b["y_mat"]), (st, w, h)
    want = json.load(open(os.path.join(GOLDEN, "enhanced_grid_digest.json")))
    for key, dig in want.items():
        w, h, st = [int(v) for v in key.split("x")]
        g = PG.enhanced_image_grid(w, h, st)
        assert hashlib.sha256(g["x_mat"].tobytes()).hexdigest() == dig[0], key
        assert hashlib.sha256(g["y_mat"].tobytes()).hexdigest() == dig[1], key


def test_flattener_matches_oracle_on_evolved_genomes():
    w, h = 48, 40
    for preset, c_dim, structure in (("circles_bw", 1, 1), ("circles", 3, 1), ("free", 3, 2)):
        grid = OG.create_grid(structure, w, h, 10)
        n_out = G.NEAT_PRESETS[preset]["num_outputs"]
        cfg = G.make_config(2, n_out)
        gc = cfg.genome_config
        n_const = 0
        for i in range(200, 260):
            g = G.synthetic_genome(preset, i, evolved=True)
            prog = G.flatten_genome(g, cfg, n_outputs=c_dim if c_dim > 1 else 1)
            n_const += any(s == G.SLOT_ONE for _, s in prog.terms)
            want = OC.render(grid, g, c_dim, w, h, gc.input_keys, gc.output_keys)
            assert np.array_equal(render_flat(prog, grid, c_dim, w, h), want), (preset, i)
        assert n_const > 5  # the constant-folding rule is actually exercised


def test_flattener_matches_oracle_on_random_topologies():
    """Seeded random feed-forward genomes (tests/fuzz_genomes.py: disabled connections, hidden chains, product aggregation,
    input-less and dangling nodes, connections leaving outputs, non-unit responses): the flattened program, interpreted
    the way render.cuh does, gives the oracle's bytes - and the oracle is byte-equal to the reference on the same genomes
    (tests/test_oracle_vs_reference.py)."""
    from fuzz_genomes import fuzz_genome
    w, h = 40, 32
    n_const = 0
    for c_dim, n_out, structure in ((1, 1, 1), (3, 3, 2), (3, 3, 1)):
        grid = OG.create_grid(structure, w, h, 10)
        cfg = G.make_config(2, n_out)
        gc = cfg.genome_config
        for seed in range(120):
            g = fuzz_genome(seed * 7 + c_dim, n_out)
            prog = G.flatten_genome(g, cfg, n_outputs=c_dim if c_dim > 1 else 1)
            n_const += any(s == G.SLOT_ONE for _, s in prog.terms)
            want = OC.render(grid, g, c_dim, w, h, gc.input_keys, gc.output_keys)
            assert np.array_equal(render_flat(prog, grid, c_dim, w, h), want), (c_dim, structure, seed)
    assert n_const > 100


def test_flattener_edge_cases():
    cfg = G.make_config(2, 1)
    g = G.Genome()
    g.nodes[0] = G.NodeGene(0, 0.25, 1.0, "sigmoid", "sum")
    prog = G.flatten_genome(g, cfg)           # output without inputs: constant bias, activation skipped
    assert len(prog.nodes) == 1 and prog.out_slots[0] & G.OUT_F32_CONST
    assert prog.terms[0][0] == float(np.float32(0.25))
    g.connections[(-1, 0)] = G.ConnectionGene((-1, 0), 2.0, enabled=False)
    assert G.flatten_genome(g, cfg).terms[0][1] == G.SLOT_ONE   # disabled connection is dropped
    g.connections[(-1, 0)].enabled = True
    g.nodes[0].aggregation = "prod"
    p = G.flatten_genome(g, cfg)
    assert p.nodes[0][1] == G.AGG_IDS["prod"] and p.terms[0] == (2.0, G.SLOT_X)
    with pytest.raises(ValueError):
        G.flatten_genome(g, G.make_config(4, 1))   # default.txt's 4 inputs (cppn.py:198 asserts too)
    blob, off, slots = G.pack_population([p, p])
    assert off[1] * 2 == off[2] == len(blob) and off[1] % 8 == 0 and slots == 4


def test_c_flattener_equals_the_python_flattener():
    """csrc/flatten.c against `flatten_genome` on every preset, plain and evolved genomes, every output count: identical
    program bytes; the genomes it hands back (constant sub-graphs needing torch's float32 transcendental kernels) take the
    Python flattener; ProgramCache.get goes through it and sees an in-place mutation."""
    if G._cflat is None:
        import __graft_entry__
        __graft_entry__.build()
        import importlib
        importlib.reload(G)
    assert G._cflat is not None
    total = handed_back = 0
    for preset in ("circles_bw", "circles", "bands", "free", "default"):
        cfg = G.make_config(2, G.NEAT_PRESETS[preset]["num_outputs"])
        gc = cfg.genome_config
        for evolved in (False, True):
            for i in range(300):
                g = G.synthetic_genome(preset, i, evolved=evolved, num_inputs=2)
                for n_out in (None, 1, 3):
                    want = G.flatten_genome(g, cfg, n_outputs=n_out)
                    r = G._cflat.flatten(g, list(gc.input_keys), list(gc.output_keys), n_out)
                    total += 1
                    if r is None:
                        handed_back += 1
                    else:
                        assert r[0] == want.to_bytes() and r[1] == want.n_slots, (preset, evolved, i, n_out)
                    fast = G.flatten_genome_fast(g, cfg, n_outputs=n_out)
                    assert fast.to_bytes() == want.to_bytes() and fast.n_slots == want.n_slots
    assert 0 < handed_back < 0.05 * total
    cfg = G.make_config(2, 3)
    cache = G.ProgramCache()
    gid, g = G.synthetic_population("circles", 4)[2]
    a = cache.get(gid, g, cfg, 3).to_bytes()
    next(iter(g.connections.values())).weight += 0.25
    b = cache.get(gid, g, cfg, 3).to_bytes()
    assert a != b and b == G.flatten_genome(g, cfg, n_outputs=3).to_bytes() and cache.fast == 2
    assert G._cflat.flatten(g, [-1, -2, -3, -4], [0, 1, 2], None) is None      # four leaves: Python raises the ValueError
    with pytest.raises(ValueError):
        G.flatten_genome_fast(g, G.make_config(4, 3))


def test_program_cache_hits_elites_and_sees_in_place_mutation(monkeypatch):
    """SURVEY §8 f row 4 (Python flattener; with the C extension `get` flattens every time): unchanged genomes re-submitted
    under the same id reuse their packed program; a genome mutated in place is re-flattened; entries of genomes that left
    the population are dropped."""
    monkeypatch.setattr(G, "_cflat", None)
    cfg = G.make_config(2, 3)
    pop = G.synthetic_population("circles", 24, evolved=True)
    cache = G.ProgramCache(keep=2)
    first = cache.flatten_population(pop, cfg, n_outputs=3)
    again = cache.flatten_population(pop, cfg, n_outputs=3)
    assert cache.misses == 24 and cache.hits == 24
    assert all(a is b for a, b in zip(first, again))
    direct = [G.flatten_genome(g, cfg, n_outputs=3).to_bytes() for _, g in pop]
    assert [p.to_bytes() for p in again] == direct
    # in-place mutation under the same id
    gid, g = pop[3]
    next(iter(g.connections.values())).weight += 0.25
    g.nodes[next(iter(g.nodes))].bias -= 0.5
    prog = cache.get(gid, g, cfg, 3)
    assert prog is not first[3] and prog.to_bytes() == G.flatten_genome(g, cfg, n_outputs=3).to_bytes()
    # another output count is another program
    assert cache.get(gid, g, cfg, 1).to_bytes() == G.flatten_genome(g, cfg, n_outputs=1).to_bytes()
    # eviction: only the survivors stay after `keep` generations
    survivors = pop[:5]
    for _ in range(3):
        cache.flatten_population(survivors, cfg, n_outputs=3)
    assert len(cache) == 5


def test_weight_layout_and_checks():
    ch = (3, 48, 96, 192)
    shapes = W.expected_shapes(160, 120, ch)
    assert shapes["predictor/ConvLSTM1/x_i1/W"] == (48, 96, 3, 3)
    assert shapes["predictor/ConvLSTM3/c_o/W"] == (1, 192, 15, 20)
    assert "predictor/ConvLSTM3/x_i1/W" not in shapes and "predictor/ConvA0/W" not in shapes
    n_params = sum(int(np.prod(s)) for s in shapes.values())
    assert abs(n_params - 8.30e6) < 0.05e6   # 8.30 M parameters
    wts = W.synthetic_weights(64, 64, (1, 4, 8, 8), seed=0)
    W.check_weights(wts, 64, 64, (1, 4, 8, 8))
    with pytest.raises(ValueError):
        W.check_weights(wts, 72, 64, (1, 4, 8, 8))   # peephole maps tie a file to one resolution


def test_shard_bounds_cover_population_in_order():
    for n in (1, 5, 32, 33, 128, 1024):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                lo, hi, per = runtime.shard_bounds(n, r, world)
                assert hi - lo <= per
                got += list(range(lo, hi))
            assert got == list(range(n))


def test_engine_cache_grows_geometrically(monkeypatch):
    made = []

    class FakeEngine:
        def __init__(self, w, h, channels, max_genomes):
            self.max_genomes, self.closed = max_genomes, False
            made.append(self)

        def load_weights(self, weights):
            self.weights = weights

        def set_conv_mode(self, mode):
            self.mode = mode

        def close(self):
            self.closed = True

    monkeypatch.setattr(runtime, "engine_factory", FakeEngine)
    monkeypatch.setattr(runtime, "_engines", {})
    a = runtime.get_engine(64, 64, (1, 4, 8, 8), "m.npz", 5)
    assert a.max_genomes == 8 and runtime.get_engine(64, 64, (1, 4, 8, 8), "m.npz", 8) is a
    b = runtime.get_engine(64, 64, (1, 4, 8, 8), "m.npz", 9)          # outgrown by one genome: 1.5x, not +1
    assert b is not a and a.closed and b.max_genomes == 12 and b.weights == "m.npz"
    assert runtime.get_engine(64, 64, (1, 4, 8, 8), "m.npz", 12) is b
    assert runtime.get_engine(64, 64, (1, 4, 8, 8), "m.npz", 40).max_genomes == 40
    assert runtime.get_engine(64, 64, (1, 4, 8, 8), "other.npz", 3) is not made[2] and len(made) == 4
    assert all(e.mode == _lib.CONV_TC and e.conv_mode == "tc" for e in made)   # EIG_CONV=auto: the tensor-core path


def test_conv_policy_of_the_drop_in_entry_points(monkeypatch):
    """get_engine picks the convolution engine (EIG_CONV=auto|tc|simt), re-applies it when the engine is
    re-created, falls back to the exact-fp32 path where the tensor-core path is missing, and after an EIG_E_RANGE."""
    class FakeEngine:
        tc_ok = True

        def __init__(self, w, h, channels, max_genomes):
            self.max_genomes, self.calls = max_genomes, 0

        def load_weights(self, weights):
            pass

        def set_conv_mode(self, mode):
            if mode == _lib.CONV_TC and not self.tc_ok:
                raise _lib.EigError(_lib.EIG_E_INVALID, "tensor-core path unavailable")
            self.mode = mode

        def close(self):
            pass

        def evaluate(self, programs, structure, render_mode, pair_mode):
            self.calls += 1
            if self.mode == _lib.CONV_TC:
                raise _lib.EigError(_lib.EIG_E_RANGE, "range")
            return np.zeros(len(programs))

    monkeypatch.setattr(runtime, "engine_factory", FakeEngine)
    monkeypatch.setattr(runtime, "_engines", {})
    monkeypatch.delenv("EIG_CONV", raising=False)
    e = runtime.get_engine(64, 64, (1, 4, 8, 8), "m.npz", 4)
    assert e.conv_mode == "tc"
    with pytest.warns(UserWarning, match="split-fp16 range"):
        out = runtime.evaluate_population(e, [object()] * 3, 1)
    assert out.shape == (3,) and e.conv_mode == "simt" and e.calls == 2
    e2 = runtime.get_engine(64, 64, (1, 4, 8, 8), "m.npz", 100)        # re-created: the fallback sticks
    assert e2 is not e and e2.conv_mode == "simt"
    monkeypatch.setattr(runtime, "_engines", {})
    monkeypatch.setenv("EIG_CONV", "simt")
    assert runtime.get_engine(64, 64, (1, 4, 8, 8), "m.npz", 4).mode == _lib.CONV_SIMT
    monkeypatch.setattr(runtime, "_engines", {})
    FakeEngine.tc_ok = False
    monkeypatch.setenv("EIG_CONV", "auto")
    assert runtime.get_engine(64, 64, (1, 4, 8, 8), "m.npz", 4).conv_mode == "simt"
    monkeypatch.setattr(runtime, "_engines", {})
    monkeypatch.setenv("EIG_CONV", "tc")
    with pytest.raises(_lib.EigError):
        runtime.get_engine(64, 64, (1, 4, 8, 8), "m.npz", 4)


def test_render_mode_follows_the_reference_branches():
    """generate_illusion.py:391,405,450: colour is a gradient only for gradient == 1, gray rounds only for gradient == 0."""
    from evolutionary_illusion_generator_b200 import engine as E
    assert [E.render_mode_for(3, g) for g in (1, 0, 2)] == [E.RENDER_GRADIENT, E.RENDER_PALETTE, E.RENDER_PALETTE]
    assert [E.render_mode_for(1, g) for g in (1, 0, 2)] == [E.RENDER_GRADIENT, E.RENDER_GRAY_ROUND, E.RENDER_GRADIENT]


def test_libeig_exports_every_symbol_of_the_header():
    header = open(os.path.join(ROOT, "include", "eig.h")).read()
    declared = set(re.findall(r"\b(eig_[a-z0-9_]+)\s*\(", header))
    assert declared == {name for name, _, _ in _lib.SYMBOLS}
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(dll, name), name
    lib = _lib.EigLibrary(_lib.LIB_PATH)
    assert lib.eig_version() >= 100


def test_product_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from evolutionary_illusion_generator_b200 import engine as E
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        E.Engine(64, 64, (1, 4, 8, 8), 2)
    lib = _lib.EigLibrary(_lib.LIB_PATH)
    ctx = ctypes.c_void_p()
    rc = lib.eig_create(ctypes.byref(ctx), 0, 64, 64, 1, (ctypes.c_int * 4)(1, 4, 8, 8), 2)
    assert rc == _lib.EIG_E_NODEVICE and b"no CPU fallback" in lib.eig_last_error()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "evolutionary_illusion_generator_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
            assert "tests.emu" not in src and "libeig_emu" not in src, f


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU oracle port timed on the host cores) runs without a GPU and prints exactly one JSON
    line with the contract's keys for the same metric / workload as the CUDA arm."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--ref-sample", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["higher_is_better"] is True
    assert d["config"]["workload"] == "c3" and d["n_gpus"] == 1 and d["value"] > 0   # the contract line is BASELINE configs[2]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "NEAT genome fitness evals/sec" in d["metric"]


def test_fold_algebra_of_the_upsampled_taps():
    """The identity behind `eig_ctx::Zf` / `build_fold_weights` (csrc/eig_api.cu): a 3x3 cross-correlation (pad 1) over a
    nearest-neighbour x2 up-sampled tensor equals, per output-pixel parity (py, px), a 2x2 cross-correlation of the
    half-resolution tensor whose taps are sums of the original ones - low-resolution offsets {-1, 0, 0} for parity 0 and
    {0, 0, +1} for parity 1 along each axis - so each parity uses 4 of the 9 low-resolution taps (`fold_tap_mask`)."""
    import torch
    import torch.nn.functional as F
    rng = np.random.RandomState(0)
    C, N, H, W = 5, 7, 6, 8
    r = torch.from_numpy(rng.randn(1, C, H, W))
    wgt = torch.from_numpy(rng.randn(N, C, 3, 3))
    up = r.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    want = F.conv2d(up, wgt, padding=1)                                   # [1, N, 2H, 2W]
    low_off = {0: (-1, 0, 0), 1: (0, 0, 1)}
    got = torch.zeros_like(want)
    for py in (0, 1):
        for px in (0, 1):
            folded = torch.zeros(N, C, 3, 3, dtype=torch.float64)
            for ky in range(3):
                for kx in range(3):
                    folded[:, :, low_off[py][ky] + 1, low_off[px][kx] + 1] += wgt[:, :, ky, kx]
            used = {(low_off[py][ky] + 1) * 3 + (low_off[px][kx] + 1) for ky in range(3) for kx in range(3)}
            assert len(used) == 4 and all(folded.reshape(N, C, 9)[:, :, t].abs().sum() == 0 for t in range(9) if t not in used)
            got[:, :, py::2, px::2] = F.conv2d(r, folded, padding=1)      # the tap-masked half-resolution convolution
    assert torch.allclose(got, want, rtol=1e-12, atol=1e-12)
