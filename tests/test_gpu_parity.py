"""Parity tests proper: the CUDA path, called through the C ABI (libeig.so), against the oracle and the committed
golden vectors.  Run on the B200 box: `pytest -m gpu`.  Nothing here reads /root/reference."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from evolutionary_illusion_generator_b200 import _lib, engine as E, genome as G, weights as W
from oracle import cppn as OC, flow as OF, grid as OG, pipeline as OPL, prednet as OP, scoring as OS

pytestmark = pytest.mark.gpu


def _sq(a, c):
    return a[..., 0] if c == 1 else a


def _programs(preset, c, idx, evolved=False):
    n_out = G.NEAT_PRESETS[preset]["num_outputs"]
    cfg = G.make_config(2, min(n_out, 3) if c == 3 else n_out)
    pop = [G.synthetic_genome(preset, i, evolved=evolved) for i in idx]
    return cfg, pop, [G.flatten_genome(g, cfg, n_outputs=c if c > 1 else 1) for g in pop]


def _pipeline_cases():
    """(npz, meta) of every whole-path golden case: plain LeCun weights and predictor-structured weights."""
    for fn in ("pipeline.npz", "pipeline_predictor.npz"):
        z = np.load(os.path.join(GOLDEN, fn))
        for m in json.loads(str(z["meta"])):
            yield z, m


def _weights_for(m):
    w, h, ch = m["w"], m["h"], tuple(m["channels"])
    if m.get("weights") == "predictor":
        return W.synthetic_predictor_weights(w, h, ch, seed=0)
    return W.synthetic_weights(w, h, ch, seed=0)


def test_loaded_library_is_the_in_tree_cuda_build(gpu_engine_factory):
    eng = gpu_engine_factory(64, 64, (1, 4, 8, 8), 2)
    assert eng.lib.path.endswith("evolutionary_illusion_generator_b200/libeig.so")
    assert eng.tdev.type == "cuda"


def test_render_matches_reference_golden_bytes(gpu_engine_factory):
    z = np.load(os.path.join(GOLDEN, "render_ref.npz"))
    for m in json.loads(str(z["meta"])):
        c, w, h = m["c_dim"], m["w"], m["h"]
        eng = gpu_engine_factory(w, h, (c, 4, 8, 8), 8)
        eng.set_grid(m["structure"])
        _, _, progs = _programs(m["preset"], c, m["indices"], m["evolved"])
        img, _ = eng.render(progs, mode=E.render_mode_for(c, m["gradient"]))
        img = img.cpu().numpy()
        for k, i in enumerate(m["indices"]):
            want = z["img_%d_%d" % (m["case"], i)]
            assert np.array_equal(_sq(img[k], c), want), (m, i, int((_sq(img[k], c) != want).sum()))


def test_render_matches_oracle_on_many_genomes(gpu_engine_factory):
    """>= 400 synthetic genomes at 160x120 (BASELINE configs C2/C3); byte-exact is the bar."""
    w, h = 160, 120
    total = bad = 0
    for preset, c in (("circles_bw", 1), ("circles", 3)):
        eng = gpu_engine_factory(w, h, (c, 4, 8, 8), 64)
        eng.set_grid(1)
        grid = OG.create_grid(1, w, h, 10)
        for start in range(0, 192, 64):
            idx = list(range(1000 + start, 1000 + start + 64))
            cfg, pop, progs = _programs(preset, c, idx, evolved=True)
            gc = cfg.genome_config
            img, x = eng.render(progs)
            img, x = img.cpu().numpy(), x.cpu().numpy()
            for k, g in enumerate(pop):
                want = OC.render(grid, g, c, w, h, gc.input_keys, gc.output_keys)
                bad += int((_sq(img[k], c) != want).sum())
                total += want.size
                assert np.array_equal(x[k], (img[k] / 255).astype(np.float32))
    print("render: %d mismatching bytes of %d" % (bad, total))
    assert bad == 0


def test_cppn_known_answers_on_gpu(gpu_engine_factory):
    cases = json.load(open(os.path.join(GOLDEN, "cppn_cases.json")))
    w, h = 64, 64
    eng = gpu_engine_factory(w, h, (1, 4, 8, 8), 2)
    for c in cases:
        g = G.Genome()
        for k, (bias, resp, act, agg) in c["nodes"].items():
            g.nodes[int(k)] = G.NodeGene(int(k), bias, resp, act, agg)
        for a, b, wgt in c["conns"]:
            g.connections[(a, b)] = G.ConnectionGene((a, b), wgt)
        scale = 1.0 / 16   # outputs land in [0,1) so the uint8 image carries the value
        eng.set_grid(grid={"x_mat": np.full((h, w), c["x"] * scale), "y_mat": np.full((h, w), c["y"] * scale)})
        img, _ = eng.render([G.flatten_genome(g, G.make_config(2, 1))])
        expect = c["expect"] if not c["conns"] else c["expect"] * scale
        assert int(img[0, 0, 0, 0]) == int(np.array([expect * 255.0]).astype(np.uint8)[0]), c["name"]


@pytest.mark.parametrize("mode", ["simt", "tc"])
def test_prednet_frames_vs_oracle(gpu_engine_factory, mode):
    for z, m in _pipeline_cases():
        name = m["name"]
        if name not in ("c2", "c3", "c2p", "c3p"):
            continue
        c, w, h, ch = m["c_dim"], m["w"], m["h"], tuple(m["channels"])
        eng = gpu_engine_factory(w, h, ch, 8)
        if mode == "tc":
            try:
                eng.set_conv_mode(_lib.CONV_TC)
            except _lib.EigError as e:
                pytest.skip("tensor-core conv not available: %s" % e)
        else:
            eng.set_conv_mode(_lib.CONV_SIMT)
        eng.set_grid(m["structure"])
        eng.load_weights(_weights_for(m))
        n = min(m["n"], 4)
        _, _, progs = _programs(m["preset"], c, list(range(n)))
        _, x = eng.render(progs)
        frames = eng.prednet(x, 20, 2).cpu().numpy()
        want = z["frames_" + name][:n]                      # [n][3][h][w(,3)]
        flips = 0
        for i in range(n):
            for k in range(3):
                d = _sq(frames[k, i], c).astype(int) - want[i, k].astype(int)
                assert np.abs(d).max() <= 1, (name, i, k)
                flips += int((d != 0).sum())
        frac = flips / want.size
        print("prednet %s %s: %d of %d bytes differ by 1 LSB (%.2e)" % (mode, name, flips, want.size, frac))
        assert frac < 1e-3


def test_flow_vs_oracle_and_cv2_golden(gpu_engine_factory):
    z = np.load(os.path.join(GOLDEN, "flow_cv2.npz"))
    for name, c in (("c2", 1), ("c3", 3)):
        keys = sorted(k[2:] for k in z.files if k.startswith("a_" + name))
        a = np.stack([z["a_" + k].reshape(120, 160, c) for k in keys])
        b = np.stack([z["b_" + k].reshape(120, 160, c) for k in keys])
        eng = gpu_engine_factory(160, 120, (c, 4, 8, 8), 8)
        corners, nc, vec, nv = [t.cpu().numpy() for t in eng.flow(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())]
        for i, k in enumerate(keys):
            assert np.array_equal(corners[i, :nc[i]], z["corners_" + k]), k          # cv2's corner list, in order
            got, cvv = vec[i, :nv[i]], z["vectors_" + k]
            assert np.array_equal(got, OF.lucas_kanade_np(_sq(a[i], c), _sq(b[i], c))), k   # bit-exact vs the oracle
            assert got.shape == cvv.shape and (len(cvv) == 0 or np.abs(got - cvv).max() <= 5e-4), k


def test_flow_random_images_all_pyramid_depths(gpu_engine_factory):
    rng = np.random.RandomState(9)
    for (h, w) in [(64, 64), (104, 112), (240, 320)]:
        base = rng.randint(0, 256, (h + 8, w + 8)).astype(np.float32)
        k = np.ones((5, 5), np.float32) / 25
        sm = sum(np.roll(np.roll(base, dy, 0), dx, 1) for dy in range(-2, 3) for dx in range(-2, 3)) / 25
        a = sm[4:-4, 4:-4].astype(np.uint8)
        b = sm[4:-4, 3:-5].astype(np.uint8)     # 1-px shift
        eng = gpu_engine_factory(w, h, (1, 4, 8, 8), 4)
        A = torch.from_numpy(np.stack([a, a])[..., None].copy()).cuda()
        Bt = torch.from_numpy(np.stack([b, a])[..., None].copy()).cuda()
        corners, nc, vec, nv = [t.cpu().numpy() for t in eng.flow(A, Bt)]
        for i, (p, q) in enumerate(((a, b), (a, a))):
            assert np.array_equal(corners[i, :nc[i]], OF.good_features(p))
            assert np.array_equal(vec[i, :nv[i]], OF.lucas_kanade_np(p, q))


def test_score_matches_reference_golden(gpu_engine_factory):
    z = np.load(os.path.join(GOLDEN, "scoring_ref.npz"))
    w, h = int(z["w"]), int(z["h"])
    eng = gpu_engine_factory(w, h, (1, 4, 8, 8), 64)
    vec = torch.from_numpy(z["vectors"]).cuda()
    nv = torch.from_numpy(z["scores"][:, 0].astype(np.int32)).cuda()
    worst = 0.0
    for st in range(4):
        got = eng.score(vec, nv, st).cpu().numpy()
        want = z["scores"][:, 1 + st]
        assert np.array_equal(np.isnan(got), np.isnan(want))
        ok = ~np.isnan(want)
        worst = max(worst, float(np.abs(got[ok] - want[ok]).max()))
    print("score: worst abs deviation from the reference's scores %.2e" % worst)
    assert worst <= 2e-6


@pytest.mark.parametrize("mode", ["simt", "tc"])
def test_whole_path_fitness_vs_oracle_golden(gpu_engine_factory, mode):
    report = []
    for z, m in _pipeline_cases():
        c, w, h, ch, n = m["c_dim"], m["w"], m["h"], tuple(m["channels"]), m["n"]
        eng = gpu_engine_factory(w, h, ch, 16)
        if mode == "tc":
            try:
                eng.set_conv_mode(_lib.CONV_TC)
            except _lib.EigError as e:
                pytest.skip("tensor-core conv not available: %s" % e)
        else:
            eng.set_conv_mode(_lib.CONV_SIMT)
        eng.set_grid(m["structure"])
        eng.load_weights(_weights_for(m))
        _, _, progs = _programs(m["preset"], c, list(range(n)))
        fit = eng.evaluate(progs, m["structure"], pair_mode=m["pair"])
        want = z["fitness_" + m["name"]]
        dbg = eng.debug_buffers(n)
        rel = np.abs(fit - want) / np.maximum(np.abs(want), 1e-12)
        rel[(fit == 0) & (want == 0)] = 0
        report.append((m["name"], float((rel <= 1e-3).mean()), rel.max(), list(dbg["nvec"]), list(z["nvec_" + m["name"]])))
    for r in report:
        print("whole path %s %s: within 1e-3: %.2f  worst rel %.2e  nvec %s vs %s" % ((mode,) + r))
    # tolerance stated by north_star: 1e-3 relative, for EVERY genome of every golden case.  (Fitness is discontinuous in
    # the uint8 frames, so a genome whose frames carry an LSB flip may land outside; such a genome would have to be named
    # here with its flip - measured on the B200: none, profiles/r2/parity_report_*.txt.)
    assert all(r[1] == 1.0 for r in report), [(r[0], r[1], r[2]) for r in report if r[1] < 1.0]


@pytest.mark.parametrize("mode", ["simt", "tc"])
def test_whole_path_fitness_vs_the_reference_itself(gpu_engine_factory, mode):
    """tests/golden/reference_pipeline.npz: fitness assigned by the REFERENCE's own get_fitnesses_neat (run unmodified
    under the Chainer shim, see tests/golden/make_golden.py) and the frames it handed to lucas_kanade.  The CUDA path
    must match it directly: fitness within 1e-3 relative (north_star), frames within 1 LSB."""
    z = np.load(os.path.join(GOLDEN, "reference_pipeline.npz"))
    worst = 0.0
    for m in json.loads(str(z["meta"])):
        w, h, ch, c = m["w"], m["h"], tuple(m["channels"]), m["c_dim"]
        cfg = G.make_config(2, G.NEAT_PRESETS[m["preset"]]["num_outputs"])
        pop = G.synthetic_population(m["preset"], m["n"], evolved=m["evolved"])
        progs = [G.flatten_genome(g, cfg, n_outputs=c if c > 1 else 1) for _, g in pop]
        eng = gpu_engine_factory(w, h, ch, m["n"])
        eng.set_conv_mode(_lib.CONV_TC if mode == "tc" else _lib.CONV_SIMT)
        eng.set_grid(m["structure"])
        eng.load_weights(W.synthetic_predictor_weights(w, h, ch, seed=m["weight_seed"]))
        fit = eng.evaluate(progs, m["structure"], E.render_mode_for(c, m["gradient"]))
        ref = z["fitness_" + m["name"]]
        frames = eng.debug_buffers(m["n"])["frames"]
        ref_frames = z["frames_" + m["name"]]
        for k in range(2):
            d = np.abs(_sq(frames[k], c).astype(int) - ref_frames[:, k].astype(int))
            assert d.max() <= 1 and (d > 0).mean() < 1e-3, (m["name"], k, int(d.max()), float((d > 0).mean()))
        assert np.allclose(fit, ref, rtol=1e-3, atol=1e-9, equal_nan=True), (m["name"], mode, fit, ref)
        nz = np.isfinite(ref) & (ref != 0)
        if nz.any():
            worst = max(worst, float(np.max(np.abs(fit[nz] / ref[nz] - 1))))
        print("%s [%s]: gpu %s reference %s" % (m["name"], mode, np.round(fit, 6), np.round(ref, 6)))
    print("GPU vs the reference's own get_fitnesses_neat [%s]: worst relative difference %.2e" % (mode, worst))


def test_edge_cases_and_errors(gpu_engine_factory):
    w, h, ch = 64, 64, (1, 4, 8, 8)
    eng = E.Engine(w, h, ch, 2)
    cfg, pop, progs = _programs("circles_bw", 1, [0, 1, 2])
    with pytest.raises(_lib.EigError) as ei:
        eng.evaluate(progs[:1], 1)
    assert ei.value.code == _lib.EIG_E_STATE                    # weights / grid not loaded
    eng.set_grid(1)
    wts = W.synthetic_weights(w, h, ch, seed=2)
    eng.load_weights(wts)
    with pytest.raises(_lib.EigError) as ei:
        eng.evaluate(progs, 1)
    assert ei.value.code == _lib.EIG_E_CAPACITY                 # population larger than max_genomes
    one = eng.evaluate(progs[:1], 2)                            # N = 1 (the reference crashes here)
    two = eng.evaluate(progs[:2], 2)
    assert one.shape == (1,) and one[0] == two[0]
    const = G.Genome()
    const.nodes[0] = G.NodeGene(0, 0.5, 1.0, "sin", "sum")       # no connections: constant image
    f = eng.evaluate([G.flatten_genome(const, cfg)], 1)
    gc = cfg.genome_config
    ref = OPL.evaluate_population([const], gc.input_keys, gc.output_keys, 1, wts, w, h, ch, 1)
    assert np.allclose(f, ref, rtol=1e-3, atol=1e-9)
    with pytest.raises(_lib.EigError):
        E.Engine(60, 64, ch, 2)                                  # w not a multiple of 8
    eng.close()


def test_full_size_properties_c2(gpu_engine_factory):
    """BASELINE configs[1]: pop 32, circles_bw, 160x120 gray.  Size-independent properties:
    determinism, shard invariance (any sub-population evaluates to the same bits), host entry == resident entry."""
    w, h, ch, n = 160, 120, (1, 16, 32, 64), 32
    eng = gpu_engine_factory(w, h, ch, n)
    eng.set_conv_mode(_lib.CONV_SIMT)
    eng.set_grid(1)
    eng.load_weights(W.synthetic_weights(w, h, ch, seed=0))
    _, _, progs = _programs("circles_bw", 1, list(range(n)))
    full = eng.evaluate(progs, 1)
    again = eng.evaluate(progs, 1)
    assert np.array_equal(full, again)
    parts = np.concatenate([eng.evaluate(progs[:16], 1), eng.evaluate(progs[16:27], 1), eng.evaluate(progs[27:], 1)])
    assert np.array_equal(full, parts)
    res = eng.upload_programs(progs)
    dev = eng.evaluate_resident(res, 1)
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy(), full)
    z = np.load(os.path.join(GOLDEN, "pipeline.npz"))
    want = z["fitness_c2"]
    rel = np.abs(full[:8] - want) / np.maximum(np.abs(want), 1e-12)
    rel[(full[:8] == 0) & (want == 0)] = 0
    print("C2 first 8 genomes rel err vs oracle:", np.array2string(rel, precision=2))
    assert (rel <= 1e-3).all(), rel
    assert (full > 0).mean() > 0.5


@pytest.mark.parametrize("name,preset,structure,w,h", [("C4", "bands", 0, 320, 240), ("C5", "free", 2, 512, 512)])
def test_largest_baseline_configs(gpu_engine_factory, name, preset, structure, w, h):
    """BASELINE configs[3] / configs[4] (320x240 Bands, 512x512 Free, colour, 3-level LK pyramid): the tensor-core path
    against the exact-fp32 SIMT path on 3 genomes, and against the CPU oracle on the first one."""
    ch, c, n = (3, 48, 96, 192), 3, 3
    wts = W.synthetic_predictor_weights(w, h, ch, seed=0)
    cfg, pop, progs = _programs(preset, c, list(range(n)))
    fits = {}
    for mode in ("simt", "tc"):
        eng = gpu_engine_factory(w, h, ch, n)
        eng.set_conv_mode(_lib.CONV_TC if mode == "tc" else _lib.CONV_SIMT)
        eng.set_grid(structure)
        eng.load_weights(wts)
        fits[mode] = eng.evaluate(progs, structure)
        assert np.array_equal(fits[mode], eng.evaluate(progs, structure))          # deterministic
        if mode == "tc":
            dbg = eng.debug_buffers(n)
    gc = cfg.genome_config
    ref, extra = OPL.evaluate_population(pop[:1], gc.input_keys, gc.output_keys, structure, wts, w, h, ch, c, keep=True)
    print("%s fitness simt %s tc %s oracle[0] %s nvec %s" % (name, fits["simt"], fits["tc"], ref, list(dbg["nvec"])))
    assert np.array_equal(dbg["image"][0], extra[0]["image"])                        # render byte-exact at full size

    def close(a, b):
        return (a == b) or abs(a - b) <= 1e-3 * max(abs(b), 1e-12) or (np.isnan(a) and np.isnan(b))
    assert close(fits["simt"][0], ref[0])
    assert all(close(a, b) for a, b in zip(fits["tc"], fits["simt"])), (fits["tc"], fits["simt"])


@pytest.mark.parametrize("c", [1, 3])
def test_every_structure_and_render_mode_end_to_end(gpu_engine_factory, c):
    """All four StructureType grids x gradient on / off (gray: np.round path, colour: palette path,
    generate_illusion.py:404-458) through the whole path on small frames, against the CPU oracle."""
    w, h, ch, n = 80, 64, (c, 4, 8, 8), 3     # Bands needs w % 10 == 0 (generate_illusion.py:222-227)
    preset = "circles_bw" if c == 1 else "circles"
    wts = W.synthetic_predictor_weights(w, h, ch, seed=5)
    cfg, pop, progs = _programs(preset, c, [3, 4, 5], evolved=True)
    gc = cfg.genome_config
    eng = gpu_engine_factory(w, h, ch, n)
    eng.load_weights(wts)
    for structure in range(4):
        eng.set_grid(structure)
        for gradient in (1, 0):
            mode = E.render_mode_for(c, gradient)
            fit = eng.evaluate(progs, structure, render_mode=mode)
            dbg = eng.debug_buffers(n)
            ref, extra = OPL.evaluate_population(pop, gc.input_keys, gc.output_keys, structure, wts, w, h, ch, c,
                                                 gradient=gradient, keep=True)
            for i in range(n):
                assert np.array_equal(_sq(dbg["image"][i], c), extra[i]["image"]), (structure, gradient, i)
            assert np.allclose(fit, ref, rtol=1e-3, atol=1e-9, equal_nan=True), (structure, gradient, fit, ref)


def test_out_of_range_activations_fail_loudly_in_tensor_core_mode(gpu_engine_factory):
    """Split-fp16 storage covers |activation| < 4094: a weight file that leaves the range must produce EIG_E_RANGE from the
    host entry point (not silent infinities); the exact-fp32 SIMT mode still evaluates it."""
    w, h, ch = 64, 64, (1, 16, 32, 64)
    wts = W.synthetic_predictor_weights(w, h, ch, seed=0)
    wts = dict(wts)
    wts["predictor/ConvA2/W"] = wts["predictor/ConvA2/W"] * np.float32(3e5)
    _, _, progs = _programs("circles_bw", 1, [0, 1])
    eng = gpu_engine_factory(w, h, ch, 2)
    eng.set_grid(1)
    eng.load_weights(wts)
    eng.set_conv_mode(_lib.CONV_TC)
    with pytest.raises(_lib.EigError) as ei:
        eng.evaluate(progs, 1)
    assert ei.value.code == _lib.EIG_E_RANGE
    eng.set_conv_mode(_lib.CONV_SIMT)
    assert eng.evaluate(progs, 1).shape == (2,)
    eng.load_weights(W.synthetic_predictor_weights(w, h, ch, seed=0))      # the flag does not stick
    eng.set_conv_mode(_lib.CONV_TC)
    assert np.all(np.isfinite(eng.evaluate(progs, 1)))


def test_streamed_evaluation_equals_host_entry_point(gpu_engine_factory):
    """Engine.evaluate_streamed (what get_fitnesses_neat runs: flatten chunk k+1 while the GPU evaluates chunk k) gives
    the bits of eig_eval_host for every chunking, repeatedly (graph replay per chunk size), and reports the range flag
    of the resident path through eig_range_check."""
    w, h, ch = 64, 64, (1, 16, 32, 64)
    cfg, pop, progs = _programs("circles_bw", 1, list(range(7)))
    items = list(enumerate(pop))
    calls = []

    def flatten(gid, g):
        calls.append(gid)
        return G.flatten_genome(g, cfg, n_outputs=1)

    eng = gpu_engine_factory(w, h, ch, 8)
    eng.set_conv_mode(_lib.CONV_TC)
    eng.set_grid(1)
    eng.load_weights(W.synthetic_predictor_weights(w, h, ch, seed=0))
    want = eng.evaluate(progs, 1)
    for chunk in (None, 3, 3, 3, 1, 7, 2):
        got = eng.evaluate_streamed(items, flatten, 1, chunk=chunk).cpu().numpy()
        assert np.array_equal(got, want, equal_nan=True), (chunk, got, want)
    assert calls[:7] == list(range(7))
    assert eng.stream_chunk(7) == 7                       # small gray genomes: one chunk
    big = gpu_engine_factory(160, 120, (3, 48, 96, 192), 8)
    assert big.stream_chunk(128) == 64 and big.stream_chunk(16) == 16
    bad = dict(W.synthetic_predictor_weights(w, h, ch, seed=0))
    bad["predictor/ConvA2/W"] = bad["predictor/ConvA2/W"] * np.float32(3e5)
    eng.load_weights(bad)
    with pytest.raises(_lib.EigError) as ei:
        eng.evaluate_streamed(items, flatten, 1, chunk=4)
    assert ei.value.code == _lib.EIG_E_RANGE
    eng.load_weights(W.synthetic_predictor_weights(w, h, ch, seed=0))      # the flag was cleared by the check
    assert np.array_equal(eng.evaluate_streamed(items, flatten, 1, chunk=4).cpu().numpy(), want, equal_nan=True)


def test_results_do_not_depend_on_what_was_evaluated_before(gpu_engine_factory):
    """A population smaller than the engine capacity must see freshly reset recurrent state in both fp16 planes of the
    split storage (the lo plane lies behind the hi plane of all `cap` genomes): same bits as a fresh, exactly sized engine,
    whatever ran before."""
    w, h, ch = 64, 64, (1, 16, 32, 64)
    _, _, progs = _programs("circles_bw", 1, list(range(8)))
    wts = W.synthetic_predictor_weights(w, h, ch, seed=0)
    for mode in (_lib.CONV_TC, _lib.CONV_SIMT):
        big = gpu_engine_factory(w, h, ch, 8)
        big.set_conv_mode(mode); big.set_grid(1); big.load_weights(wts)
        big.evaluate(progs, 1)                                     # leaves state of 8 other genomes behind
        for n in (3, 1, 5):
            fresh = gpu_engine_factory(w, h, ch, n)
            fresh.set_conv_mode(mode); fresh.set_grid(1); fresh.load_weights(wts)
            want = fresh.evaluate(progs[:n], 1)
            assert np.array_equal(big.evaluate(progs[:n], 1), want, equal_nan=True), (mode, n)
            big.evaluate(progs[3:8], 1)


def test_graph_replay_is_transparent(gpu_engine_factory):
    """The library replays everything after the render from a CUDA graph from the third identical call on.  Direct run,
    capture run and replays must give the same bits; a new weight file, another population size and another structure
    must not see stale graphs."""
    w, h, ch = 64, 64, (1, 16, 32, 64)
    _, _, progs = _programs("circles_bw", 1, [0, 1, 2, 3])
    w_a, w_b = W.synthetic_predictor_weights(w, h, ch, seed=1), W.synthetic_predictor_weights(w, h, ch, seed=2)
    eng = gpu_engine_factory(w, h, ch, 4)
    eng.set_conv_mode(_lib.CONV_TC)
    eng.set_grid(1)
    eng.load_weights(w_a)
    runs = [eng.evaluate(progs, 1) for _ in range(4)]          # direct, capture, replay, replay
    assert all(np.array_equal(runs[0], r) for r in runs[1:])
    three = [eng.evaluate(progs[:3], 1) for _ in range(3)]     # other n: its own graph
    assert all(np.array_equal(three[0], r) for r in three) and np.array_equal(three[0], runs[0][:3])
    eng.load_weights(w_b)                                      # drops the graphs
    b_runs = [eng.evaluate(progs, 1) for _ in range(3)]
    fresh = gpu_engine_factory(w, h, ch, 4)
    fresh.set_conv_mode(_lib.CONV_TC); fresh.set_grid(1); fresh.load_weights(w_b)
    want_b = fresh.evaluate(progs, 1)
    assert all(np.array_equal(want_b, r) for r in b_runs) and not np.array_equal(want_b, runs[0])
    eng.set_grid(2)                                            # other structure: new grid planes + its own graph
    free_runs = [eng.evaluate(progs, 2) for _ in range(3)]
    fresh.set_grid(2)
    assert all(np.array_equal(fresh.evaluate(progs, 2), r) for r in free_runs)


def test_tensor_core_path_is_the_one_that_runs(gpu_engine_factory):
    """In tensor-core mode every layer-1..3 convolution of the BASELINE networks is a tcgen05 launch (no silent fall back
    to the SIMT kernel): counted with the library's per-class launch instrumentation."""
    import ctypes as C
    for ch, c, preset in (((1, 16, 32, 64), 1, "circles_bw"), ((3, 48, 96, 192), 3, "circles")):
        eng = gpu_engine_factory(160, 120, ch, 4)
        eng.set_conv_mode(_lib.CONV_TC)
        eng.set_grid(1)
        eng.load_weights(W.synthetic_predictor_weights(160, 120, ch, seed=0))
        _, _, progs = _programs(preset, c, [0, 1, 2])
        ms, cnt = (C.c_double * 8)(), (C.c_int64 * 8)()
        eng.lib.check(eng.lib.eig_profile_begin(eng.ctx))
        eng.evaluate(progs, 1)
        eng.lib.check(eng.lib.eig_profile_end(eng.ctx, ms, cnt))
        n_tc, n_simt = int(cnt[2]), int(cnt[1])
        print("channels %s: %d tcgen05 conv launches, %d SIMT conv launches" % (ch, n_tc, n_simt))
        assert n_simt == 0
        assert n_tc == 21 * (9 if ch[1] >= 32 else 8) - 2  # A2 A3 L3 L2 L1 P1+Z P2 P3 (+ ConvA1 for wide first layers); no P2 / P3 on the last step


def test_reference_call_surface(tmp_path):
    """The drop-in modules with the reference's names and arguments: get_fitnesses_neat (generate_illusion.py:478-673),
    get_vectors / calculate_fitness (fitness_calculator.py:468-548), lucas_kanade (optical_flow.py:40-89)."""
    from PIL import Image
    from evolutionary_illusion_generator_b200 import fitness_calculator as FC, generate_illusion as GI, optical_flow as OFL
    w, h, ch, n = 64, 64, (1, 4, 8, 8), 4
    wts = W.synthetic_predictor_weights(w, h, ch, seed=3)
    model = str(tmp_path / "prednet.npz")
    W.save_npz(model, wts)
    cfg, pop, _ = _programs("circles_bw", 1, list(range(n)))
    population = [(100 + i, g) for i, g in enumerate(pop)]
    best_dir = str(tmp_path / "best")
    assert GI.get_fitnesses_neat(GI.StructureType.Free, population, model, cfg, w, h, ch, c_dim=1, best_dir=best_dir) is None
    t0 = __import__("time").perf_counter()
    GI.get_fitnesses_neat(GI.StructureType.Free, population, model, cfg, w, h, ch, c_dim=1, best_dir=best_dir, export_async=True)
    t_async = __import__("time").perf_counter() - t0
    GI.wait_for_exports()
    print("get_fitnesses_neat with background export: %.1f ms" % (1e3 * t_async))
    assert GI.program_cache.hits + GI.program_cache.fast >= n                     # the second generation re-used every flattened program
    gc = cfg.genome_config
    ref, extra = OPL.evaluate_population(pop, gc.input_keys, gc.output_keys, 2, wts, w, h, ch, 1, keep=True)
    got = np.array([g.fitness for _, g in population])
    assert all(isinstance(g.fitness, float) for _, g in population)
    assert np.allclose(got, ref, rtol=1e-3, atol=1e-9, equal_nan=True), (got, ref)
    for name, size in (("best.png", (w, h)), ("best_black_bg.png", (w, h)), ("best_flow.png", (w, h)), ("enhanced.png", (800, 800))):
        assert Image.open(os.path.join(best_dir, name)).size == size, name
    best, best_i = 0, 0
    for i, f in enumerate(ref):           # generate_illusion.py:625-628: `>=` keeps the last of equal scores, NaN never wins
        if f >= best:
            best, best_i = f, i
    assert np.array_equal(np.asarray(Image.open(os.path.join(best_dir, "best.png"))), extra[best_i]["image"])
    # single-image rating path: input image vs extension #2
    img_path = str(tmp_path / "img.png")
    Image.fromarray(extra[0]["image"], "L").save(img_path)
    vec = FC.get_vectors(img_path, model, ch, w, h)
    want_vec = OF.lucas_kanade_np(extra[0]["image"], extra[0]["frames"][2])
    if len(want_vec) == 0:
        assert len(vec) == 1 and vec[0] is None
    else:
        assert np.allclose(np.asarray(vec), want_vec, atol=2e-3)
        f = FC.calculate_fitness(GI.StructureType.Free, vec, img_path, w, h)
        assert np.isclose(f, OS.fitness_from_vectors(2, want_vec, w, h), rtol=1e-3, atol=1e-9, equal_nan=True)
    assert FC.calculate_fitness(GI.StructureType.Circles, [None], img_path, w, h) == 0.0
    # file-based lucas_kanade: overlay + csv written, vectors identical to the oracle's
    f1, f2 = str(tmp_path / "a.png"), str(tmp_path / "b.png")
    Image.fromarray(extra[1]["frames"][0], "L").save(f1)
    Image.fromarray(extra[1]["frames"][1], "L").save(f2)
    res = OFL.lucas_kanade(f1, f2, output_path=str(tmp_path / "flow"), verbose=0)
    want = OF.lucas_kanade_np(extra[1]["frames"][0], extra[1]["frames"][1])
    assert np.array_equal(np.asarray(res["vectors"], np.float32).reshape(-1, 4), want.reshape(-1, 4))
    assert os.path.isfile(str(tmp_path / "flow" / "a.png")) and os.path.isfile(str(tmp_path / "flow" / "csv" / "a.csv"))


def test_exported_files_vs_the_reference_itself(tmp_path):
    """get_fitnesses_neat of the drop-in package on the populations of tests/golden/reference_pipeline.npz: genome.fitness
    within 1e-3 of what the reference assigned, and best.png / best_black_bg.png / enhanced.png (800x800 mosaic) pixel-identical
    to the files the reference wrote for its best genome (generate_illusion.py:650-671)."""
    from PIL import Image
    from evolutionary_illusion_generator_b200 import generate_illusion as GI
    z = np.load(os.path.join(GOLDEN, "reference_pipeline.npz"))
    for m in json.loads(str(z["meta"])):
        if "export_best_" + m["name"] not in z.files:
            continue
        w, h, ch, c = m["w"], m["h"], tuple(m["channels"]), m["c_dim"]
        model = str(tmp_path / (m["name"] + ".npz"))
        W.save_npz(model, W.synthetic_predictor_weights(w, h, ch, seed=m["weight_seed"]))
        cfg = G.make_config(2, G.NEAT_PRESETS[m["preset"]]["num_outputs"])
        pop = G.synthetic_population(m["preset"], m["n"], evolved=m["evolved"])
        best_dir = str(tmp_path / ("best_" + m["name"]))
        GI.get_fitnesses_neat(GI.StructureType(m["structure"]), pop, model, cfg, w, h, ch, c_dim=c, best_dir=best_dir,
                              gradient=m["gradient"])
        got = np.array([g.fitness for _, g in pop])
        assert np.allclose(got, z["fitness_" + m["name"]], rtol=1e-3, atol=1e-9, equal_nan=True)
        for k in ("best", "best_black_bg", "enhanced"):
            mine = np.asarray(Image.open(os.path.join(best_dir, k + ".png")))
            ref = z["export_%s_%s" % (k, m["name"])]
            assert mine.shape == ref.shape and np.array_equal(mine, ref), (m["name"], k, mine.shape, ref.shape)
        assert np.asarray(Image.open(os.path.join(best_dir, "best_flow.png"))).shape == z["export_best_flow_" + m["name"]].shape
        print("%s: best.png, best_black_bg.png, enhanced.png identical to the reference's files" % m["name"])


def test_test_prednet_mirror_vs_the_reference_frames(gpu_engine_factory, tmp_path, monkeypatch):
    """`call_prednet.test_prednet` of the drop-in package with the argument list `get_fitnesses_neat` uses
    (generate_illusion.py:531-535): same file names, frames within 1 LSB of the PNGs the reference's own test_prednet wrote
    (tests/golden/reference_pipeline.npz), state carried frame to frame and reset after every extension block."""
    from PIL import Image
    from evolutionary_illusion_generator_b200 import call_prednet as CP
    monkeypatch.chdir(tmp_path)                       # test_log.txt lands in the working directory, like the reference's
    z = np.load(os.path.join(GOLDEN, "reference_pipeline.npz"))
    for m in json.loads(str(z["meta"])):
        if m["name"] not in ("r_small_free", "r_small_colour_palette"):
            continue
        w, h, ch, c, n = m["w"], m["h"], tuple(m["channels"]), m["c_dim"], m["n"]
        model = str(tmp_path / (m["name"] + ".npz"))
        W.save_npz(model, W.synthetic_predictor_weights(w, h, ch, seed=m["weight_seed"]))
        cfg = G.make_config(2, G.NEAT_PRESETS[m["preset"]]["num_outputs"])
        pop = G.synthetic_population(m["preset"], n, evolved=m["evolved"])
        eng = gpu_engine_factory(w, h, ch, n)
        eng.set_grid(m["structure"])
        img, _ = eng.render([G.flatten_genome(g, cfg, n_outputs=c if c > 1 else 1) for _, g in pop],
                            mode=E.render_mode_for(c, m["gradient"]))
        img = img.cpu().numpy()
        out = tmp_path / ("pred_" + m["name"])
        out.mkdir()
        repeated = []
        for i in range(n):
            path = str(tmp_path / ("%s_%d.png" % (m["name"], i)))
            (Image.fromarray(img[i]) if c == 3 else Image.fromarray(img[i][:, :, 0], "L")).save(path)
            repeated += [path] * 20
        CP.test_prednet(initmodel=model, sequence_list=[repeated], size=[w, h], channels=list(ch), gpu=0,
                        output_dir=str(out), skip_save_frames=1, extension_start=20, extension_duration=2,
                        reset_at=22, verbose=0, c_dim=c)
        assert len(list(out.iterdir())) == 22 * n
        ref = z["frames_" + m["name"]]
        for i in range(n):
            got = [np.asarray(Image.open(out / ("%010d.png" % (20 * i + 19)))),
                   np.asarray(Image.open(out / ("%010d_extended.png" % (20 * i + 20))))]
            for k in range(2):
                d = np.abs(got[k].astype(int) - ref[i, k].astype(int))
                assert d.max() <= 1 and (d > 0).mean() < 1e-3, (m["name"], i, k, int(d.max()), float((d > 0).mean()))
        assert len(open("test_log.txt").read().splitlines()) == 20 * n - 1
        print("%s: %d files, frames within 1 LSB of the reference's test_prednet output" % (m["name"], 22 * n))


def test_single_image_rating_vs_the_reference_itself(tmp_path):
    """`fitness_calculator.get_vectors` / `calculate_fitness` of the drop-in package against what the reference's own
    functions returned for the same image and weight file (tests/golden/reference_single_image.npz)."""
    from PIL import Image
    from evolutionary_illusion_generator_b200 import fitness_calculator as FC, generate_illusion as GI
    z = np.load(os.path.join(GOLDEN, "reference_single_image.npz"))
    for m in json.loads(str(z["meta"])):
        w, h, ch, c = m["w"], m["h"], tuple(m["channels"]), m["c_dim"]
        model = str(tmp_path / (m["name"] + ".npz"))
        W.save_npz(model, W.synthetic_predictor_weights(w, h, ch, seed=m["weight_seed"]))
        for i in range(len(m["genomes"])):
            img = z["images_" + m["name"]][i]
            path = str(tmp_path / ("%s_%d.png" % (m["name"], i)))
            (Image.fromarray(img) if c == 3 else Image.fromarray(img, "L")).save(path)
            vec = FC.get_vectors(path, model, ch, w, h)
            n = int(z["nvec_" + m["name"]][i])
            assert (len(vec) == n) if n else (vec[0] is None)
            assert np.allclose(np.asarray(vec, np.float32).reshape(-1, 4), z["vectors_" + m["name"]][i, :n], atol=2e-3)
            for k, st in enumerate((GI.StructureType.Circles, GI.StructureType.Free)):
                ref = z["scores_" + m["name"]][i, k]
                got = FC.calculate_fitness(st, vec, path, w, h)
                if np.isnan(ref):
                    assert got == 0.0                      # the reference raises UnboundLocalError here
                else:
                    assert np.isclose(got, ref, rtol=1e-3, atol=1e-9), (m["name"], i, st, got, ref)
            print("%s image %d: %d vectors, scores match the reference %s" % (m["name"], i, n, np.round(z["scores_" + m["name"]][i], 5)))


def test_tcgen05_conv_self_check():
    """tests/gpu/tc_check (built by csrc/build.sh): the tcgen05 conv against a float64 CPU convolution and, epilogue by
    epilogue, against the exact-fp32 SIMT kernel on identical inputs."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gpu", "tc_check")
    assert os.path.isfile(exe), "tests/gpu/tc_check missing: run __graft_entry__.build()"
    r = subprocess.run([exe, "16"], capture_output=True, text=True, timeout=300)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-500:]
