"""The real kernel sources (csrc/*.cuh) executed by the TEST-ONLY fiber emulator (tests/emu/cuda_emu.h) and
compared with the oracle.  This is how indexing / barriers / shuffles are checked in the GPU-less build
container; the B200 run of the same comparisons is tests/test_gpu_parity.py."""
import numpy as np
import torch

from evolutionary_illusion_generator_b200 import _lib as _lib_mod, engine as E, genome as G, weights as W
from oracle import cppn as OC, flow as OF, grid as OG, pipeline as OPL, prednet as OP, scoring as OS


def _squeeze(a, c):
    return a[..., 0] if c == 1 else a


def test_render_kernel_bytes(emu_lib):
    w, h = 64, 56
    for preset, ch, gradient in (("circles_bw", (1, 4, 8, 8), 1), ("circles_bw", (1, 4, 8, 8), 0),
                                 ("circles", (3, 4, 8, 8), 1), ("circles", (3, 4, 8, 8), 0)):
        c = ch[0]
        eng = E.Engine(w, h, ch, 8, lib=emu_lib)
        eng.set_grid(1)
        cfg = G.make_config(2, G.NEAT_PRESETS[preset]["num_outputs"])
        gc = cfg.genome_config
        pop = [G.synthetic_genome(preset, i, evolved=bool(i % 2)) for i in range(6)]
        const = G.Genome()
        const.nodes = {k: G.NodeGene(k, 0.3 + 0.2 * k, 1.0, "sigmoid", "sum") for k in gc.output_keys}
        pop.append(const)
        progs = [G.flatten_genome(g, cfg, n_outputs=c) for g in pop]
        img, x = eng.render(progs, mode=E.render_mode_for(c, gradient))
        grid = OG.create_grid(1, w, h, 10)
        for i, g in enumerate(pop):
            want = OC.render(grid, g, c, w, h, gc.input_keys, gc.output_keys, gradient=gradient)
            assert np.array_equal(_squeeze(img[i].numpy(), c), want), (preset, gradient, i)
            assert np.array_equal(_squeeze(x[i].numpy(), c), (want / 255).astype(np.float32))
        eng.close()


def test_render_kernel_on_random_topologies(emu_lib):
    """cppn_render_kernel (host-compiled) on seeded random genome topologies (tests/fuzz_genomes.py) = the oracle's bytes."""
    from fuzz_genomes import fuzz_genome
    w, h = 64, 56
    for c, n_out, structure, gradient in ((1, 1, 1, 1), (3, 3, 2, 1), (3, 3, 1, 0), (1, 1, 2, 0)):
        eng = E.Engine(w, h, (c, 4, 8, 8), 8, lib=emu_lib)
        eng.set_grid(structure)
        cfg = G.make_config(2, n_out)
        gc = cfg.genome_config
        grid = OG.create_grid(structure, w, h, 10)
        for base in range(0, 24, 8):
            pop = [fuzz_genome(1000 + (base + i) * 7 + c, n_out) for i in range(8)]
            img, _ = eng.render([G.flatten_genome(g, cfg, n_outputs=c) for g in pop], mode=E.render_mode_for(c, gradient))
            for i, g in enumerate(pop):
                want = OC.render(grid, g, c, w, h, gc.input_keys, gc.output_keys, gradient=gradient)
                assert np.array_equal(_squeeze(img[i].numpy(), c), want), (c, structure, gradient, base + i)
        eng.close()


def test_prednet_kernels_frames(emu_lib):
    w, h = 64, 56   # layer sizes 56/28/14/7: partial tiles and an odd top layer
    for preset, ch in (("circles_bw", (1, 4, 8, 12)), ("circles", (3, 6, 8, 20))):
        c = ch[0]
        eng = E.Engine(w, h, ch, 4, lib=emu_lib)
        eng.set_grid(1)
        wts = W.synthetic_weights(w, h, ch, seed=1, bias_std=0.1)
        eng.load_weights(wts)
        cfg = G.make_config(2, G.NEAT_PRESETS[preset]["num_outputs"])
        pop = [G.synthetic_genome(preset, i) for i in range(2)]
        img, x = eng.render([G.flatten_genome(g, cfg, n_outputs=c) for g in pop])
        frames = eng.prednet(x, n_input_steps=3, n_ext=2)
        net = OP.PredNetOracle(wts, ch, w, h)
        for i in range(len(pop)):
            want = OP.run_genome_frames(net, _squeeze(img[i].numpy(), c), repeat=3, extension=2)
            for k in range(3):
                d = _squeeze(frames[k, i].numpy(), c).astype(int) - want[k].astype(int)
                assert np.abs(d).max() <= 1 and (d != 0).mean() < 2e-3, (preset, i, k)
        eng.close()


def test_stepping_api_and_test_image_list_mirror(emu_lib, tmp_path):
    """eig_prednet_reset / eig_prednet_forward (state carried from call to call) and the `call_prednet.test_image_list`
    mirror built on them: frames written for 3 repeats + 2 extensions, twice in a row (state reset after the extension
    block), equal the whole-sequence kernel path; a sequence of DISTINCT frames equals the oracle stepped by hand."""
    from PIL import Image
    from evolutionary_illusion_generator_b200 import call_prednet as CP
    w, h, ch = 64, 56, (1, 4, 8, 12)
    eng = E.Engine(w, h, ch, 4, lib=emu_lib)
    eng.set_grid(1)
    wts = W.synthetic_weights(w, h, ch, seed=1, bias_std=0.1)
    eng.load_weights(wts)
    cfg = G.make_config(2, 1)
    pop = [G.synthetic_genome("circles_bw", i) for i in range(2)]
    img, x = eng.render([G.flatten_genome(g, cfg, n_outputs=1) for g in pop])
    frames = eng.prednet(x, n_input_steps=3, n_ext=2).numpy()
    paths = []
    for i in range(2):
        paths.append(str(tmp_path / ("in_%d.png" % i)))
        Image.fromarray(img[i].numpy()[:, :, 0], "L").save(paths[-1])
    out = tmp_path / "pred"
    out.mkdir()
    with open(tmp_path / "log.txt", "w") as logf:
        step = CP.test_image_list(eng, [paths[0]] * 3 + [paths[1]] * 3, None, str(out), ch, [w, h], [0, 0], 0, logf,
                                  skip_save_frames=1, extension_start=3, extension_duration=2, verbose=0, c=1)
    assert step == 6
    for i in range(2):      # generate_illusion.py:543-546 naming: prediction i*3+2, extensions (i*3+3)+j
        got = [np.asarray(Image.open(out / ("%010d.png" % (i * 3 + 2))))] + \
              [np.asarray(Image.open(out / ("%010d_extended.png" % (i * 3 + 3 + j)))) for j in range(2)]
        for k in range(3):
            assert np.array_equal(got[k], frames[k, i, :, :, 0]), (i, k)
    assert len((tmp_path / "log.txt").read_text().splitlines()) == 5          # no loss line for the last frame
    # distinct frames, no extension: against the oracle network stepped by hand
    net = OP.PredNetOracle(wts, ch, w, h)
    eng.prednet_reset(1)
    for k in range(4):
        frame_in = img[k % 2].numpy()
        pred, u8 = eng.prednet_forward(x[k % 2:k % 2 + 1])
        with torch.no_grad():
            p0 = net.step(torch.from_numpy(OP.image_to_input(_squeeze(frame_in, 1)))[None])
        want = OP.prediction_to_image(p0[0].numpy())
        d = u8[0, :, :, 0].numpy().astype(int) - np.asarray(want).reshape(h, w).astype(int)
        assert np.abs(d).max() <= 1 and (d != 0).mean() < 2e-3, k
    eng.close()


def test_flow_and_score_kernels(emu_lib):
    rng = np.random.RandomState(3)
    import cv2
    w, h = 112, 104   # two pyramid levels
    a = cv2.GaussianBlur(rng.randint(0, 256, (h, w, 3)).astype(np.uint8), (7, 7), 2)
    b = cv2.warpAffine(a, np.float32([[1, 0.003, 0.2], [-0.003, 1, -0.1]]), (w, h), borderMode=cv2.BORDER_REFLECT)
    flat = np.full((h, w, 3), 128, np.uint8)
    eng = E.Engine(w, h, (3, 4, 8, 8), 4, lib=emu_lib)
    A = torch.from_numpy(np.stack([a, flat, a]))
    Bt = torch.from_numpy(np.stack([b, flat, a]))
    corners, nc, vec, nv = eng.flow(A, Bt)
    for i, (p, q) in enumerate(((a, b), (flat, flat), (a, a))):
        want_c = OF.good_features(OF.to_gray(p))
        want_v = OF.lucas_kanade_np(p, q)
        assert np.array_equal(corners[i, :nc[i]].numpy(), want_c)
        assert np.array_equal(vec[i, :nv[i]].numpy(), want_v)
        for st in range(4):
            got = float(eng.score(vec[i:i + 1], nv[i:i + 1], st)[0])
            want = OS.fitness_from_vectors(st, want_v, w, h)
            assert (np.isnan(got) and np.isnan(want)) or abs(got - want) <= 1e-6 * max(1.0, abs(want)), (i, st)
    assert int(nc[1]) == 0 and int(nv[1]) == 0   # flat image: no corners, sentinel path, fitness 0
    eng.close()


def test_whole_path_matches_oracle(emu_lib):
    for (w, h, preset, ch, structure, pair) in [(64, 64, "circles_bw", (1, 4, 8, 8), 1, 0),
                                                (64, 64, "free", (3, 4, 6, 8), 2, 1)]:
        c, n = ch[0], 3
        eng = E.Engine(w, h, ch, n, lib=emu_lib)
        eng.set_grid(structure)
        wts = W.synthetic_weights(w, h, ch, seed=2)
        eng.load_weights(wts)
        cfg = G.make_config(2, G.NEAT_PRESETS[preset]["num_outputs"])
        gc = cfg.genome_config
        pop = [G.synthetic_genome(preset, i) for i in range(n)]
        progs = [G.flatten_genome(g, cfg, n_outputs=c) for g in pop]
        fit = eng.evaluate(progs, structure, pair_mode=pair)
        ref, ex = OPL.evaluate_population(pop, gc.input_keys, gc.output_keys, structure, wts, w, h, ch, c,
                                          pair_mode=pair, keep=True)
        dbg = eng.debug_buffers(n)
        assert list(dbg["nvec"]) == [len(e["vectors"]) for e in ex]
        assert np.allclose(fit, ref, rtol=1e-3, atol=1e-9)
        eng.close()


def test_raw_ctypes_binding_of_integration_md(emu_lib):
    """The binding INTEGRATION.md section 3 shows (plain ctypes, no helper classes), run against the host-compiled library:
    eig_create -> eig_load_weights -> eig_set_grid -> eig_eval_host, same fitness as the Engine wrapper."""
    import ctypes as C
    from conftest import EMU_SO
    from evolutionary_illusion_generator_b200.grid import create_grid
    lib = C.CDLL(EMU_SO)
    lib.eig_last_error.restype = C.c_char_p
    w, h, channels, n = 64, 64, (1, 4, 8, 8), 3
    ctx = C.c_void_p()
    ch = (C.c_int * 4)(*channels)
    assert lib.eig_create(C.byref(ctx), 0, w, h, channels[0], ch, n) == 0, lib.eig_last_error()
    z = W.synthetic_predictor_weights(w, h, channels, seed=3)
    names = sorted(z)
    arrs = [np.ascontiguousarray(z[k], np.float32) for k in names]
    shp = np.ones((len(names), 4), np.int64)
    for i, a in enumerate(arrs):
        shp[i, :a.ndim] = a.shape
    assert lib.eig_load_weights(ctx, len(names), (C.c_char_p * len(names))(*[k.encode() for k in names]),
                                (C.c_void_p * len(names))(*[a.ctypes.data for a in arrs]),
                                shp.ctypes.data_as(C.c_void_p)) == 0, lib.eig_last_error()
    g = create_grid(1, w, h, 10)
    x = np.ascontiguousarray(g["x_mat"], np.float64)
    y = np.ascontiguousarray(g["y_mat"], np.float64)
    assert lib.eig_set_grid(ctx, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p)) == 0
    cfg = G.make_config(2, 1)
    pop = [G.synthetic_genome("circles_bw", i) for i in range(n)]
    blob, offsets, max_slots = G.pack_population([G.flatten_genome(p, cfg, n_outputs=1) for p in pop])
    fit = np.empty(n, np.float64)
    rc = lib.eig_eval_host(ctx, blob.ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(C.c_void_p), n, max_slots, 1, 0, 0,
                           fit.ctypes.data_as(C.c_void_p))
    assert rc == 0, lib.eig_last_error()
    lib.eig_destroy(ctx)
    eng = E.Engine(w, h, channels, n, lib=emu_lib)
    eng.set_grid(1)
    eng.load_weights(z)
    want = eng.evaluate([G.flatten_genome(p, cfg, n_outputs=1) for p in pop], 1)
    eng.close()
    assert np.array_equal(fit, want, equal_nan=True) and np.any(fit > 0)


def test_get_fitnesses_neat_on_the_host_compiled_library_vs_the_reference(emu_lib, tmp_path, monkeypatch):
    """The drop-in `get_fitnesses_neat` with the kernel sources compiled for the host (exact-fp32 path) against what the
    REFERENCE's own get_fitnesses_neat assigned to the same population and weight file (tests/golden/reference_pipeline.npz):
    the CPU-side proof of the end-to-end parity the GPU tests repeat on the B200."""
    import json
    import os
    from PIL import Image
    from conftest import GOLDEN
    from evolutionary_illusion_generator_b200 import generate_illusion as GI, runtime
    monkeypatch.setattr(runtime, "engine_factory", lambda w, h, ch, n, **kw: E.Engine(w, h, ch, n, lib=emu_lib, **kw))
    monkeypatch.setattr(runtime, "_engines", {})
    monkeypatch.setattr(GI, "_render_engines", {})
    monkeypatch.setattr(GI, "ENHANCED_SIZE", 64)            # the 800x800 mosaic is slow in the emulator
    monkeypatch.setattr(GI, "program_cache", G.ProgramCache())
    z = np.load(os.path.join(GOLDEN, "reference_pipeline.npz"))
    m = [m for m in json.loads(str(z["meta"])) if m["name"] == "r_small_gray"][0]
    w, h, ch, c = m["w"], m["h"], tuple(m["channels"]), m["c_dim"]
    model = str(tmp_path / "model.npz")
    W.save_npz(model, W.synthetic_predictor_weights(w, h, ch, seed=m["weight_seed"]))
    cfg = G.make_config(2, 1)
    pop = G.synthetic_population(m["preset"], m["n"], evolved=m["evolved"])
    best_dir = str(tmp_path / "best")
    GI.get_fitnesses_neat(GI.StructureType(m["structure"]), pop, model, cfg, w, h, ch, c_dim=c, best_dir=best_dir,
                          gradient=m["gradient"])
    got = np.array([g.fitness for _, g in pop])
    ref = z["fitness_" + m["name"]]
    assert all(isinstance(g.fitness, float) for _, g in pop)
    assert np.allclose(got, ref, rtol=1e-3, atol=1e-9), (got, ref)
    grid = OG.create_grid(m["structure"], w, h, 10)
    gc = cfg.genome_config
    want = OC.render(grid, pop[int(np.argmax(ref))][1], c, w, h, gc.input_keys, gc.output_keys)
    assert np.array_equal(np.asarray(Image.open(os.path.join(best_dir, "best.png"))), want)
    assert sorted(os.listdir(best_dir)) == ["best.png", "best_black_bg.png", "best_flow.png", "enhanced.png"]
    GI.get_fitnesses_neat(GI.StructureType(m["structure"]), pop, model, cfg, w, h, ch, c_dim=c, best_dir=best_dir,
                          gradient=m["gradient"], export_best=False)
    assert GI.program_cache.hits + GI.program_cache.fast >= m["n"] and np.array_equal(got, np.array([g.fitness for _, g in pop]))
    for eng in list(runtime._engines.values()) + list(GI._render_engines.values()):
        eng.close()


def test_single_image_rating_on_the_host_compiled_library_vs_the_reference(emu_lib, tmp_path, monkeypatch):
    """`fitness_calculator.get_vectors` / `calculate_fitness` of the drop-in package, kernels compiled for the host, against
    what the reference's own two functions returned for the same image and weights (reference_single_image.npz)."""
    import json
    import os
    from PIL import Image
    from conftest import GOLDEN
    from evolutionary_illusion_generator_b200 import fitness_calculator as FC, generate_illusion as GI, runtime
    monkeypatch.setattr(runtime, "engine_factory", lambda w, h, ch, n, **kw: E.Engine(w, h, ch, n, lib=emu_lib, **kw))
    monkeypatch.setattr(runtime, "_engines", {})
    z = np.load(os.path.join(GOLDEN, "reference_single_image.npz"))
    m = json.loads(str(z["meta"]))[0]
    w, h, ch = m["w"], m["h"], tuple(m["channels"])
    model = str(tmp_path / "model.npz")
    W.save_npz(model, W.synthetic_predictor_weights(w, h, ch, seed=m["weight_seed"]))
    path = str(tmp_path / "image.png")
    Image.fromarray(z["images_" + m["name"]][0], "L").save(path)
    vec = FC.get_vectors(path, model, ch, w, h)
    n = int(z["nvec_" + m["name"]][0])
    assert len(vec) == n and np.allclose(np.asarray(vec).reshape(-1, 4), z["vectors_" + m["name"]][0, :n], atol=2e-3)
    for k, st in enumerate((GI.StructureType.Circles, GI.StructureType.Free)):
        assert np.isclose(FC.calculate_fitness(st, vec, path, w, h), z["scores_" + m["name"]][0, k], rtol=1e-3)
    assert FC.calculate_fitness(GI.StructureType.Circles, [None], path, w, h) == 0.0
    for eng in runtime._engines.values():
        eng.close()


def test_error_text_and_options_belong_to_the_context(emu_lib):
    """two contexts in one process (the main engine and the render-only engine of enhanced.png) keep
    their own error text; eig_set_option validates its keys and masks."""
    a = E.Engine(64, 64, (1, 4, 8, 8), 2, lib=emu_lib)
    b = E.Engine(64, 64, (1, 4, 8, 8), 2, lib=emu_lib)
    assert emu_lib.eig_set_option(a.ctx, b"passes.L7", 7) == _lib_mod.EIG_E_INVALID
    assert b"unknown key passes.L7" in emu_lib.eig_error(a.ctx)
    assert emu_lib.eig_set_option(b.ctx, b"passes.all", 0) == _lib_mod.EIG_E_INVALID
    assert b"empty pass mask" in emu_lib.eig_error(b.ctx)
    assert b"unknown key" in emu_lib.eig_error(a.ctx)                 # untouched by b's failure
    assert b"empty pass mask" in emu_lib.eig_last_error()              # the thread's most recent message
    for key, v in ((b"passes.all", 7), (b"passes.L", 6), (b"passes.A2", 5), (b"early_until", 3), (b"early_mask", 4), (b"graphs", 0)):
        assert emu_lib.eig_set_option(a.ctx, key, v) == 0, key
    a.close(); b.close()


def test_render_only_context(emu_lib):
    """eig_create_render: the CPPN stage alone, for `get_image_from_cppn` and the enhanced mosaic - any image
    size (here 30 x 22, not a multiple of 8 and below the LK window), bytes equal to the oracle, everything else refused."""
    w, h = 30, 22
    eng = E.Engine(w, h, (3, 4, 4, 4), 2, lib=emu_lib, render_only=True)
    grid = OG.create_grid(2, w, h, 10)
    eng.set_grid(grid=grid)
    cfg = G.make_config(2, 3)
    pop = [G.synthetic_genome("circles", i, evolved=True) for i in (3, 4)]
    img, x = eng.render([G.flatten_genome(g, cfg, n_outputs=3) for g in pop])
    gc = cfg.genome_config
    for k, g in enumerate(pop):
        assert np.array_equal(img[k].numpy(), OC.render(grid, g, 3, w, h, gc.input_keys, gc.output_keys))
    assert emu_lib.eig_prednet_reset(eng.ctx, 1, None) == _lib_mod.EIG_E_STATE
    assert b"only renders" in emu_lib.eig_error(eng.ctx)
    eng.close()
