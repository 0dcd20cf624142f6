"""Pins the oracle against the REFERENCE's own code, imported unmodified under stubs (build container only;
skipped on the GPU box where /root/reference does not exist)."""
import numpy as np
import pytest

from conftest import ref_available
from evolutionary_illusion_generator_b200 import genome as G, grid as PG
from oracle import cppn as OC, grid as OG, scoring as OS

pytestmark = [pytest.mark.needs_reference,
              pytest.mark.skipif(not ref_available(), reason="reference tree not mounted")]


@pytest.fixture(scope="module")
def ns():
    import ref_harness
    return ref_harness.load()


def test_grids_match_reference(ns):
    for s in range(4):
        for (w, h) in [(160, 120), (64, 64)]:
            if s == 0 and w % 10:
                continue
            ref = ns.gi.create_grid(ns.gi.StructureType(s), w, h, 10)
            for impl in (OG.create_grid(s, w, h, 10), PG.create_grid(s, w, h, 10)):
                assert np.array_equal(np.asarray(ref["x_mat"]).reshape(h, w), impl["x_mat"])
                assert np.array_equal(np.asarray(ref["y_mat"]).reshape(h, w), impl["y_mat"])


def test_enhanced_grid_matches_reference(ns):
    for s, w, h in ((1, 120, 120), (3, 90, 120), (0, 60, 60), (2, 75, 75)):
        ref = ns.gi.enhanced_image_grid(w, h, ns.gi.StructureType(s))
        for impl in (OG.enhanced_image_grid(w, h, s), PG.enhanced_image_grid(w, h, s)):
            assert np.array_equal(ref["x_mat"], impl["x_mat"]) and np.array_equal(ref["y_mat"], impl["y_mat"])


def test_render_matches_reference_on_fresh_genomes(ns):
    w, h = 64, 64
    grid = ns.gi.create_grid(ns.gi.StructureType.Circles, w, h, 10)
    for preset, c_dim in (("circles_bw", 1), ("circles", 3)):
        cfg = G.make_config(2, G.NEAT_PRESETS[preset]["num_outputs"])
        gc = cfg.genome_config
        for i in range(100, 130):
            g = G.synthetic_genome(preset, i, evolved=bool(i % 2))
            ref = np.asarray(ns.gi.get_image_from_cppn(grid, g, c_dim, w, h, cfg))
            got = OC.render(grid, g, c_dim, w, h, gc.input_keys, gc.output_keys)
            assert np.array_equal(ref, got), (preset, i)


def test_render_matches_reference_on_random_topologies(ns):
    from fuzz_genomes import fuzz_genome
    w, h = 40, 32
    for c_dim, n_out, structure in ((1, 1, 1), (3, 3, 2), (3, 3, 1)):
        grid = ns.gi.create_grid(ns.gi.StructureType(structure), w, h, 10)
        cfg = G.make_config(2, n_out)
        gc = cfg.genome_config
        for seed in range(60):
            g = fuzz_genome(seed * 7 + c_dim, n_out)
            ref = np.asarray(ns.gi.get_image_from_cppn(grid, g, c_dim, w, h, cfg))
            got = OC.render(grid, g, c_dim, w, h, gc.input_keys, gc.output_keys)
            assert np.array_equal(ref, got.reshape(ref.shape)), (c_dim, structure, seed)


def test_scoring_matches_reference_functions(ns):
    rng = np.random.RandomState(11)
    w, h = 160, 120
    for _ in range(40):
        n = rng.randint(2, 100)
        v = np.zeros((n, 4), np.float32)
        v[:, 0], v[:, 1] = rng.uniform(0, w, n), rng.uniform(0, h, n)
        v[:, 2:] = rng.normal(0, 0.1, (n, 2))
        assert OS.strength_number(v, 0.3) == ns.fc.strength_number(v, 0.3)
        assert OS.swarm_score(v) == ns.fc.swarm_score(v)
        assert OS.rotation_symmetry_score(v, w, h, [0, h / 2]) == ns.fc.rotation_symmetry_score(v, w, h, [0, h / 2])
        assert OS.horizontal_symmetry_score(v, [0, 60.0]) == ns.fc.horizontal_symmetry_score(v, [0, 60.0])


def test_reference_test_cppn_cases_pass_under_stub():
    """The reference's own four known-answer tests (pytorch_neat/tests/test_cppn.py) run here unmodified
    (in a subprocess: they import the inner package as top-level `pytorch_neat`)."""
    import subprocess
    import sys
    code = (
        "import sys, types\n"
        "sys.path.insert(0, '/root/repo/tests/golden'); import ref_harness\n"
        "neat = types.ModuleType('neat'); g = types.ModuleType('neat.graphs')\n"
        "g.required_for_output = ref_harness.required_for_output; neat.graphs = g\n"
        "sys.modules['neat'] = neat; sys.modules['neat.graphs'] = g\n"
        "sys.path.insert(0, '/root/reference/pytorch_neat'); sys.path.insert(0, '/root/reference/pytorch_neat/tests')\n"
        "import test_cppn as t\n"
        "t.test_cppn_simple(); t.test_cppn_unconnected(); t.test_cppn_call(); t.test_cppn_deep_call(); print('OK4')\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "OK4" in out.stdout, out.stderr[-2000:]


def test_reference_get_fitnesses_neat_runs_and_reproduces_the_fixture(ns, tmp_path):
    """The smallest case of tests/golden/reference_pipeline.npz, re-run live: the reference's whole fitness function,
    unmodified, under the Chainer shim.  Also checks that the parameter paths of the reference's own model definition
    (net.py:128-157, L.Classifier) are exactly the keys `weights.py` writes."""
    import json
    import os
    import sys
    from conftest import GOLDEN
    sys.path.insert(0, GOLDEN)
    import make_golden
    import chainer
    from evolutionary_illusion_generator_b200 import weights as W
    from chainer_prednet.PredNet import net
    model = chainer.links.Classifier(net.PredNet(64, 64, (1, 4, 8, 8)))
    assert sorted(p.lstrip("/") for p, _ in model.namedparams()) == sorted(W.synthetic_weights(64, 64, (1, 4, 8, 8), seed=0))
    z = np.load(os.path.join(GOLDEN, "reference_pipeline.npz"))
    case = make_golden.REFERENCE_CASES[0]
    fit, frames, _ = make_golden.run_reference_case(ns, case, str(tmp_path))
    assert np.array_equal(fit, z["fitness_" + case[0]], equal_nan=True)
    assert np.array_equal(frames, z["frames_" + case[0]])
    assert json.loads(str(z["meta"]))[0]["name"] == case[0]


def test_lucas_kanade_mirror_vs_the_reference_function(ns, emu_lib, tmp_path, monkeypatch):
    """`optical_flow.lucas_kanade` of the drop-in package (kernels compiled for the host) against the reference's own
    function on the reference's sample pair optical_flow/penguin{1,2}.png and on an evolved 160x120 image pair."""
    import os
    import ref_harness
    from evolutionary_illusion_generator_b200 import engine as E, optical_flow as OFL, runtime
    monkeypatch.setattr(runtime, "engine_factory", lambda w, h, ch, n: E.Engine(w, h, ch, n, lib=emu_lib))
    monkeypatch.setattr(OFL, "_flow_engines", {})
    ref_dir = os.path.join(ref_harness.REF, "optical_flow")
    pairs = [(os.path.join(ref_dir, "penguin1.png"), os.path.join(ref_dir, "penguin2.png"))]
    small = os.path.join(ref_harness.REF, "illusions_rating", "EIGEN-images")
    imgs = sorted(os.path.join(r, f) for r, _, fs in os.walk(small) for f in fs if f == "small.png")
    if len(imgs) >= 2:
        pairs.append((imgs[0], imgs[1]))
    for f1, f2 in pairs:
        want = ns.of.lucas_kanade(f1, f2, str(tmp_path / "ref"), save=False, verbose=0)["vectors"]
        got = OFL.lucas_kanade(f1, f2, output_path=str(tmp_path / "mine"), save=True, verbose=0)
        assert len(got["vectors"]) == len(want), (f1, len(got["vectors"]), len(want))
        if want:
            assert np.allclose(np.asarray(got["vectors"], np.float32), np.asarray(want, np.float32), atol=2e-3), f1
        stem = os.path.splitext(os.path.basename(f1))[0]
        assert os.path.isfile(tmp_path / "mine" / (stem + ".png")) and os.path.isfile(tmp_path / "mine" / "csv" / (stem + ".csv"))
    for eng in OFL._flow_engines.values():
        eng.close()


def test_get_image_from_cppn_mirror_vs_the_reference_function(ns, emu_lib):
    """`generate_illusion.get_image_from_cppn` of the drop-in package (render kernel compiled for the host) returns the very
    PIL image the reference's function returns: every structure it can render, gray / colour, with and without gradient,
    white and black background, on synthetic and fuzzed genomes."""
    from fuzz_genomes import fuzz_genome
    from evolutionary_illusion_generator_b200 import engine as E, generate_illusion as GI
    w, h = 64, 56
    for c_dim, n_out, preset in ((1, 1, "circles_bw"), (3, 3, "circles")):
        eng = E.Engine(w, h, (c_dim, 4, 8, 8), 8, lib=emu_lib)
        cfg = G.make_config(2, n_out)
        genomes = [G.synthetic_genome(preset, i, evolved=bool(i % 2)) for i in range(4)] + \
                  [fuzz_genome(500 + i, n_out) for i in range(4)]
        for structure in (ns.gi.StructureType.Circles, ns.gi.StructureType.Free, ns.gi.StructureType.CirclesFree):
            grid = ns.gi.create_grid(structure, w, h, 10)
            for k, g in enumerate(genomes):
                gradient, bg = k % 2, (k // 2) % 2
                ref = ns.gi.get_image_from_cppn(grid, g, c_dim, w, h, cfg, bg=bg, gradient=gradient)
                got = GI.get_image_from_cppn(grid, g, c_dim, w, h, cfg, bg=bg, gradient=gradient, engine=eng)
                assert got.mode == ref.mode and got.size == ref.size
                assert np.array_equal(np.asarray(got), np.asarray(ref)), (c_dim, int(structure), k)
        eng.close()


def test_test_prednet_mirror_vs_the_reference_on_distinct_frames(ns, emu_lib, tmp_path, monkeypatch):
    """A sequence of DISTINCT frames with an extension block in the middle, through the reference's own `test_prednet`
    (Chainer stand-in) and through the drop-in `call_prednet.test_prednet` (kernels compiled for the host): same file
    names, every frame within 1 LSB, same number of loss lines."""
    import os
    from PIL import Image
    from evolutionary_illusion_generator_b200 import call_prednet as CP, engine as E, runtime, weights as W
    monkeypatch.setattr(runtime, "engine_factory", lambda w, h, ch, n: E.Engine(w, h, ch, n, lib=emu_lib))
    monkeypatch.setattr(runtime, "_engines", {})
    monkeypatch.chdir(tmp_path)
    w, h, ch = 64, 64, (1, 4, 8, 8)
    model = str(tmp_path / "model.npz")
    W.save_npz(model, W.synthetic_predictor_weights(w, h, ch, seed=3))
    cfg = G.make_config(2, 1)
    grid = OG.create_grid(1, w, h, 10)
    gc = cfg.genome_config
    paths = []
    for i in range(4):
        img = OC.render(grid, G.synthetic_genome("circles_bw", i), 1, w, h, gc.input_keys, gc.output_keys)
        paths.append(str(tmp_path / ("frame_%d.png" % i)))
        Image.fromarray(img, "L").save(paths[-1])
    sequence = [paths[0], paths[1], paths[2], paths[3], paths[1], paths[0]]      # extension after frame 3, then two more
    kw = dict(size=[w, h], channels=list(ch), skip_save_frames=1, extension_start=3, extension_duration=2, reset_at=5,
              verbose=0, c_dim=1)
    os.makedirs("ref_out")
    os.makedirs("my_out")
    ns.call_prednet.test_prednet(initmodel=model, sequence_list=[sequence], gpu=-1, output_dir="ref_out", **kw)
    ref_log = open("test_log.txt").read().splitlines()
    CP.test_prednet(initmodel=model, sequence_list=[sequence], gpu=0, output_dir="my_out", **kw)
    my_log = open("test_log.txt").read().splitlines()
    assert sorted(os.listdir("my_out")) == sorted(os.listdir("ref_out")) and len(os.listdir("ref_out")) == 6 + 2 * 2
    for name in os.listdir("ref_out"):
        a = np.asarray(Image.open(os.path.join("ref_out", name))).astype(int)
        b = np.asarray(Image.open(os.path.join("my_out", name))).astype(int)
        assert np.abs(a - b).max() <= 1 and (a != b).mean() < 2e-3, name
    assert len(my_log) == len(ref_log) == 5
    for mine, ref in zip(my_log, ref_log):
        assert mine.split(",")[0] == ref.split(",")[0]
        assert np.isclose(float(mine.split(",")[1]), float(ref.split(",")[1]), rtol=1e-3, atol=1e-7)
    for eng in runtime._engines.values():
        eng.close()
