"""Generates the committed golden fixtures.  Run here (build container), where /root/reference exists:

    python tests/golden/make_golden.py

  render_ref.npz   : images produced by the REFERENCE's own `get_image_from_cppn` (imported unmodified under the
                     stubs of ref_harness.py) for seeded synthetic genomes  -> pins oracle/cppn.py + flattener + K1
  scoring_ref.npz  : scores of the REFERENCE's `calculate_fitness` on seeded random vector sets -> pins
                     oracle/scoring.py + K9
  cppn_cases.json  : the four known answers of the reference's pytorch_neat/tests/test_cppn.py:27-92
  flow_cv2.npz     : frame pairs (produced by the oracle PredNet on oracle renders) with the output of the
                     cv2 4.13 calls the reference makes (corners, vectors) -> pins oracle/flow.py + K5-K8
  pipeline.npz     : oracle fitness for seeded populations/weights at 64x64 and 160x120 -> whole-path anchor
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness  # noqa: E402
from evolutionary_illusion_generator_b200 import genome as G, weights as W  # noqa: E402
from oracle import cppn as OC, flow as OF, grid as OG, pipeline as OPL, prednet as OP  # noqa: E402

RENDER_CASES = [  # (preset, c_dim, w, h, structure, gradient, evolved, indices)
    ("circles_bw", 1, 160, 120, 1, 1, False, list(range(0, 6))),
    ("circles_bw", 1, 160, 120, 1, 0, True, list(range(6, 10))),
    ("circles", 3, 160, 120, 1, 1, True, list(range(0, 6))),
    ("circles", 3, 64, 64, 1, 0, False, list(range(6, 9))),
    ("free", 3, 64, 64, 2, 1, True, list(range(0, 4))),
    ("bands", 3, 160, 120, 0, 1, False, list(range(0, 3))),
]


def make_render(ns):
    out = {}
    meta = []
    for ci, (preset, c_dim, w, h, structure, gradient, evolved, idx) in enumerate(RENDER_CASES):
        n_out = G.NEAT_PRESETS[preset]["num_outputs"]
        # the reference crashes on 6-output colour configs and on Bands (SURVEY.md "defects"): run it on the
        # documented fix-ups (first 3 outputs, planes reshaped to (h,w))
        cfg = G.make_config(2, min(n_out, 3) if c_dim == 3 else n_out)
        grid = ns.gi.create_grid(ns.gi.StructureType(structure), w, h, 10)
        grid = {k: np.asarray(v).reshape(h, w) for k, v in grid.items()}
        for i in idx:
            g = G.synthetic_genome(preset, i, evolved=evolved)
            if c_dim == 3 and n_out > 3:
                g.nodes = {k: v for k, v in g.nodes.items()}
            img = np.asarray(ns.gi.get_image_from_cppn(grid, g, c_dim, w, h, cfg, bg=1, gradient=gradient))
            out["img_%d_%d" % (ci, i)] = img
        meta.append(dict(case=ci, preset=preset, c_dim=c_dim, w=w, h=h, structure=structure, gradient=gradient,
                         evolved=evolved, indices=idx))
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "render_ref.npz"), **out)
    print("render_ref.npz:", len(out) - 1, "images")


def make_scoring(ns):
    rng = np.random.RandomState(7)
    w, h = 160, 120
    sets, scores = [], []
    for trial in range(60):
        n = int(rng.randint(1, 100))
        v = np.zeros((n, 4), np.float32)
        v[:, 0] = rng.uniform(0, w, n)
        v[:, 1] = rng.uniform(0, h, n)
        v[:, 2:] = rng.normal(0, rng.choice([0.03, 0.1, 0.25]), (n, 2))
        if trial % 5 == 0:
            v[:, :2] = np.round(v[:, :2])
        row = []
        for st in (0, 1, 2, 3):
            try:
                s = ns.fc.calculate_fitness(ns.fc.StructureType(st), v, "x.png", w, h)
            except UnboundLocalError:
                s = 0.0
            row.append(float(s))
        pad = np.zeros((100, 4), np.float32)
        pad[:n] = v
        sets.append(pad)
        scores.append([n] + row)
    np.savez_compressed(os.path.join(HERE, "scoring_ref.npz"), vectors=np.stack(sets), scores=np.array(scores), w=w, h=h)
    print("scoring_ref.npz:", len(sets), "vector sets")


def make_cppn_cases():
    # pytorch_neat/tests/test_cppn.py:27-92, restated as genomes: (nodes, connections, inputs, expected)
    cases = [
        dict(name="simple", nodes={"0": [0.0, 1.0, "identity", "sum"]}, conns=[[-1, 0, 1.0]], x=3.0, y=0.0, expect=3.0),
        dict(name="unconnected", nodes={"0": [0.5, 1.0, "identity", "sum"]}, conns=[], x=3.0, y=0.0, expect=0.5),
        dict(name="call_b", nodes={"0": [0.0, 1.0, "identity", "sum"]}, conns=[[-1, 0, 1.0], [-2, 0, 1.0]], x=1.5, y=2.0,
             expect=3.5),
        dict(name="deep_call_b", nodes={"0": [0.0, 1.0, "identity", "sum"], "1": [0.0, 1.0, "identity", "sum"]},
             conns=[[-2, 1, 1.0], [-1, 0, 1.0], [1, 0, 1.0]], x=1.5, y=2.0, expect=3.5, n_outputs=1),
    ]
    json.dump(cases, open(os.path.join(HERE, "cppn_cases.json"), "w"), indent=1)
    print("cppn_cases.json:", len(cases))


def make_flow_and_pipeline():
    import torch
    torch.set_num_threads(8)
    out_flow, out_pipe = {}, {}
    cases = [("c2", "circles_bw", 1, (1, 16, 32, 64), 160, 120, 1, 0, 8),
             ("c1", "circles_bw", 1, (1, 16, 32, 64), 64, 64, 1, 1, 3),
             ("c3", "circles", 3, (3, 48, 96, 192), 160, 120, 1, 0, 3),
             ("small_free", "free", 3, (3, 6, 8, 12), 64, 64, 2, 0, 4)]
    meta = []
    for name, preset, c_dim, ch, w, h, structure, pair, n in cases:
        wts = W.synthetic_weights(w, h, ch, seed=0)
        n_out = G.NEAT_PRESETS[preset]["num_outputs"]
        cfg = G.make_config(2, n_out)
        pop = [G.synthetic_genome(preset, i) for i in range(n)]
        fit, ex = OPL.evaluate_population(pop, cfg.genome_config.input_keys, cfg.genome_config.output_keys, structure,
                                          wts, w, h, ch, c_dim, pair_mode=pair, keep=True)
        out_pipe["fitness_" + name] = fit
        out_pipe["nvec_" + name] = np.array([len(e["vectors"]) for e in ex])
        out_pipe["frames_" + name] = np.stack([np.stack(e["frames"]) for e in ex])
        meta.append(dict(name=name, preset=preset, c_dim=c_dim, channels=ch, w=w, h=h, structure=structure, pair=pair, n=n))
        if name in ("c2", "c3"):
            for i, e in enumerate(ex[:4]):
                a, b = e["frames"][0], e["frames"][1]
                cvv = OF.lucas_kanade_cv2(a, b)
                import cv2
                gray = OF.to_gray(a)
                cc = cv2.goodFeaturesToTrack(gray, 100, 0.3, 7, blockSize=7)
                cc = np.zeros((0, 2), np.float32) if cc is None else cc.reshape(-1, 2)
                key = "%s_%d" % (name, i)
                out_flow["a_" + key], out_flow["b_" + key] = a, b
                out_flow["corners_" + key], out_flow["vectors_" + key] = cc, cvv
        print(name, "fitness", np.round(fit, 5))
    out_pipe["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "pipeline.npz"), **out_pipe)
    np.savez_compressed(os.path.join(HERE, "flow_cv2.npz"), **out_flow)
    print("pipeline.npz, flow_cv2.npz written")


def make_enhanced_digest(ns):
    """sha256 of the planes the reference's own `enhanced_image_grid` returns (generate_illusion.py:121-193)."""
    import hashlib
    out = {}
    for w, h, st in ((800, 800, 1), (800, 800, 0), (300, 240, 3)):
        g = ns.gi.enhanced_image_grid(w, h, ns.gi.StructureType(st))
        out["%dx%dx%d" % (w, h, st)] = [hashlib.sha256(np.ascontiguousarray(g["x_mat"], np.float64).tobytes()).hexdigest(),
                                         hashlib.sha256(np.ascontiguousarray(g["y_mat"], np.float64).tobytes()).hexdigest()]
    json.dump(out, open(os.path.join(HERE, "enhanced_grid_digest.json"), "w"), indent=1)
    print("enhanced_grid_digest.json:", list(out))


def make_pipeline_predictor():
    """Same as the pipeline cases above but with `synthetic_predictor_weights` (P0 tracks the frame, flow of a
    few hundredths of a pixel, non-zero fitness for most genomes): oracle only, /root/reference not needed."""
    import torch
    torch.set_num_threads(8)
    out, meta = {}, []
    cases = [("c2p", "circles_bw", 1, (1, 16, 32, 64), 160, 120, 1, 0, 12),
             ("c3p", "circles", 3, (3, 48, 96, 192), 160, 120, 1, 0, 6),
             ("c1p", "circles_bw", 1, (1, 16, 32, 64), 64, 64, 1, 1, 3)]
    for name, preset, c_dim, ch, w, h, structure, pair, n in cases:
        wts = W.synthetic_predictor_weights(w, h, ch, seed=0)
        cfg = G.make_config(2, G.NEAT_PRESETS[preset]["num_outputs"])
        pop = [G.synthetic_genome(preset, i) for i in range(n)]
        fit, ex = OPL.evaluate_population(pop, cfg.genome_config.input_keys, cfg.genome_config.output_keys, structure,
                                          wts, w, h, ch, c_dim, pair_mode=pair, keep=True)
        out["fitness_" + name] = fit
        out["nvec_" + name] = np.array([len(e["vectors"]) for e in ex])
        out["frames_" + name] = np.stack([np.stack(e["frames"]) for e in ex])
        meta.append(dict(name=name, preset=preset, c_dim=c_dim, channels=ch, w=w, h=h, structure=structure, pair=pair, n=n,
                         weights="predictor"))
        print(name, "fitness", np.round(fit, 5), "nvec", out["nvec_" + name])
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "pipeline_predictor.npz"), **out)


REFERENCE_CASES = [
    # name, preset, c_dim, channels, w, h, structure, gradient, n genomes, evolved, weight seed
    ("r_small_gray", "circles_bw", 1, (1, 4, 8, 8), 64, 64, 1, 1, 6, True, 3),
    ("r_small_free", "circles_bw", 1, (1, 16, 32, 64), 64, 64, 2, 1, 6, True, 0),
    ("r_small_colour_palette", "circles", 3, (3, 6, 12, 24), 64, 64, 1, 0, 4, True, 5),
    ("r_circlesfree", "circles_bw", 1, (1, 16, 32, 64), 160, 120, 3, 1, 3, False, 0),
    ("r_small_gray_round", "circles_bw", 1, (1, 4, 8, 8), 64, 64, 1, 0, 4, False, 3),
    ("r_c2", "circles_bw", 1, (1, 16, 32, 64), 160, 120, 1, 1, 8, False, 0),
    ("r_c3", "circles", 3, (3, 48, 96, 192), 160, 120, 1, 1, 4, False, 0),
    ("r_320x240", "circles", 3, (3, 48, 96, 192), 320, 240, 1, 1, 2, False, 0),      # three LK pyramid levels
]


# BASELINE configs[4] shape (512x512 colour, Free scoring branch, generate_illusion.py:590-607) through the reference
# itself: `python make_golden.py reference-large` -> reference_pipeline_large.npz.  3-output genomes: the reference's
# colour render indexes image_array[:, :, c] for every output node, so the 6-output free config overflows its 3-channel
# array (SURVEY.md "defects").  configs[3] (Bands) cannot be recorded this way: the reference's own get_fitnesses_neat
# raises IndexError on the (1, w*h) Bands planes in get_image_from_cppn (generate_illusion.py:400; same defect list) -
# its 320x240 resolution is covered by r_320x240 above and the Bands scoring branch by scoring_ref.npz.
REFERENCE_LARGE_CASES = [
    ("r_c5_free", "circles", 3, (3, 48, 96, 192), 512, 512, 2, 1, 2, False, 0),
]

EXPORT_CASES = ("r_small_free", "r_small_colour_palette")


def run_reference_case(ns, case, workdir):
    """The reference's own `get_fitnesses_neat` (generate_illusion.py:478-673), UNMODIFIED, on a seeded population and a
    seeded weight file, with Chainer replaced by tests/golden/chainer_shim and "the GPU" being numpy.  Returns the
    fitness the reference assigned to every genome plus the two frames it fed to lucas_kanade for every genome."""
    from PIL import Image
    name, preset, c_dim, ch, w, h, structure, gradient, n, evolved, wseed = case
    os.makedirs(os.path.join(workdir, "best"), exist_ok=True)
    model = os.path.join(workdir, "model.npz")
    W.save_npz(model, W.synthetic_predictor_weights(w, h, ch, seed=wseed))
    cfg = G.make_config(2, G.NEAT_PRESETS[preset]["num_outputs"])    # circles / circles_bw: 3 / 1 outputs
    pop = G.synthetic_population(preset, n, evolved=evolved)
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        ns.gi.get_fitnesses_neat(ns.gi.StructureType(structure), pop, model, cfg, w, h, list(ch), c_dim=c_dim,
                                 best_dir=os.path.join(workdir, "best"), gradient=gradient)
        repeat, ext = 20, 2
        frames = []
        for i in range(n):              # generate_illusion.py:543-546
            i0 = i * repeat + repeat - 1
            f0 = np.asarray(Image.open("temp/prediction/%s.png" % str(i0).zfill(10)))
            f1 = np.asarray(Image.open("temp/prediction/%s_extended.png" % str(i0 + ext - 1).zfill(10)))
            frames.append(np.stack([f0, f1]))
        exported = {k: np.asarray(Image.open(os.path.join(workdir, "best", k + ".png")))
                    for k in ("best", "best_black_bg", "best_flow", "enhanced")}      # generate_illusion.py:650-671
    finally:
        os.chdir(cwd)
    return np.array([float(g.fitness) for _, g in pop]), np.stack(frames), exported


def make_reference_pipeline(ns, cases=None, fname="reference_pipeline.npz"):
    """tests/golden/reference_pipeline.npz: fitness vectors and frames produced by the reference itself."""
    import tempfile
    out, meta = {}, []
    for case in (cases or REFERENCE_CASES):
        with tempfile.TemporaryDirectory() as d:
            fit, frames, exported = run_reference_case(ns, case, d)
        name = case[0]
        out["fitness_" + name], out["frames_" + name] = fit, frames
        if name in EXPORT_CASES:        # the files the reference writes for the best genome of the generation
            for k, v in exported.items():
                out["export_%s_%s" % (k, name)] = v
        meta.append(dict(zip(("name", "preset", "c_dim", "channels", "w", "h", "structure", "gradient", "n", "evolved",
                              "weight_seed"), case)))
        print(name, "reference fitness", np.round(fit, 6))
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, fname), **out)


def make_full_populations():
    """full_populations.npz: the oracle on EVERY genome of the two 160x120 bench populations (BASELINE configs[1] pop 32,
    configs[2] pop 128; bench.build_population, predictor weights seed 0): fitness, number of flow vectors and the two
    frames handed to the flow stage for the first 8 genomes.  ~3 minutes on 8 cores; /root/reference not needed."""
    import torch
    torch.set_num_threads(8)
    out, meta = {}, []
    for name, preset, c_dim, ch, w, h, structure, n in (("c2", "circles_bw", 1, (1, 16, 32, 64), 160, 120, 1, 32),
                                                        ("c3", "circles", 3, (3, 48, 96, 192), 160, 120, 1, 128)):
        wts = W.synthetic_predictor_weights(w, h, ch, seed=0)
        cfg = G.make_config(2, G.NEAT_PRESETS[preset]["num_outputs"])
        pop = [G.synthetic_genome(preset, i) for i in range(n)]
        fit, ex = OPL.evaluate_population(pop, cfg.genome_config.input_keys, cfg.genome_config.output_keys, structure,
                                          wts, w, h, ch, c_dim, keep=True, flow_impl="cv2")
        out["fitness_" + name] = fit
        out["nvec_" + name] = np.array([len(e["vectors"]) for e in ex])
        out["frames_" + name] = np.stack([np.stack(e["frames"][:2]) for e in ex[:8]])
        meta.append(dict(name=name, preset=preset, c_dim=c_dim, channels=ch, w=w, h=h, structure=structure, n=n,
                         weights="predictor", weight_seed=0))
        print(name, "nonzero fitness", int((fit > 0).sum()), "of", n, "mean nvec", out["nvec_" + name].mean())
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "full_populations.npz"), **out)


SINGLE_IMAGE_CASES = [
    # name, preset, c_dim, channels, w, h, genome indices, weight seed
    ("s_small_gray", "circles_bw", 1, (1, 16, 32, 64), 64, 64, [0, 1, 2], 0),
    ("s_c3", "circles", 3, (3, 48, 96, 192), 160, 120, [1, 2], 0),
]


def run_single_image_case(ns, case, workdir):
    """The reference's single-image rating flow (fitness_calculator.py:468-548: `get_vectors` pairs the input image with
    extension #2, `calculate_fitness` scores it), unmodified, on images rendered by the reference's own
    `get_image_from_cppn`.  A `calculate_fitness` call that hits the reference's UnboundLocalError is recorded as NaN."""
    name, preset, c_dim, ch, w, h, idx, wseed = case
    model = os.path.join(workdir, "model.npz")
    W.save_npz(model, W.synthetic_predictor_weights(w, h, ch, seed=wseed))
    cfg = G.make_config(2, G.NEAT_PRESETS[preset]["num_outputs"])
    grid = ns.gi.create_grid(ns.gi.StructureType.Circles, w, h, 10)
    cwd = os.getcwd()
    os.chdir(workdir)
    images, vectors, scores = [], [], []
    try:
        for i in idx:
            img = ns.gi.get_image_from_cppn(grid, G.synthetic_genome(preset, i), c_dim, w, h, cfg)
            path = os.path.join(workdir, "img_%d.png" % i)
            img.save(path, "PNG")
            vec = ns.fc.get_vectors(path, model, list(ch), w, h)
            row = []
            for st in (ns.gi.StructureType.Circles, ns.gi.StructureType.Free):
                try:
                    row.append(float(ns.fc.calculate_fitness(st, vec, path, w, h)) if vec[0] is not None else float("nan"))
                except UnboundLocalError:
                    row.append(float("nan"))
            images.append(np.asarray(img))
            vectors.append(np.zeros((0, 4), np.float32) if vec[0] is None else np.asarray(vec, np.float32).reshape(-1, 4))
            scores.append(row)
    finally:
        os.chdir(cwd)
    return images, vectors, np.array(scores)


def make_reference_single_image(ns):
    import tempfile
    out, meta = {}, []
    for case in SINGLE_IMAGE_CASES:
        with tempfile.TemporaryDirectory() as d:
            images, vectors, scores = run_single_image_case(ns, case, d)
        name = case[0]
        out["images_" + name] = np.stack(images)
        out["nvec_" + name] = np.array([len(v) for v in vectors])
        pad = np.zeros((len(vectors), 100, 4), np.float32)
        for i, v in enumerate(vectors):
            pad[i, :len(v)] = v
        out["vectors_" + name], out["scores_" + name] = pad, scores
        meta.append(dict(zip(("name", "preset", "c_dim", "channels", "w", "h", "genomes", "weight_seed"), case)))
        print(name, "vectors", out["nvec_" + name], "scores (Circles, Free)", np.round(scores, 5).tolist())
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "reference_single_image.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "reference":
        make_reference_pipeline(ref_harness.load())
        make_reference_single_image(ref_harness.load())
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "reference-large":
        make_reference_pipeline(ref_harness.load(), REFERENCE_LARGE_CASES, "reference_pipeline_large.npz")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "full":
        make_full_populations()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "single":
        make_reference_single_image(ref_harness.load())
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "predictor":
        make_pipeline_predictor()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "enhanced":
        make_enhanced_digest(ref_harness.load())
        sys.exit(0)
    ns = ref_harness.load()
    make_render(ns)
    make_scoring(ns)
    make_cppn_cases()
    make_enhanced_digest(ns)
    make_flow_and_pipeline()
    make_pipeline_predictor()
    make_reference_pipeline(ns)
    make_reference_single_image(ns)
