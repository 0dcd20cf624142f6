"""Drop-in for the fitness path of /root/reference/generate_illusion.py, backed by libeig.so on a B200.

Same names, argument meaning and side effects on `genome.fitness` as the reference:
  StructureType                  generate_illusion.py:25-29
  create_grid                    generate_illusion.py:196-317   (grid.py)
  get_image_from_cppn            generate_illusion.py:372-460   -> PIL image, rendered by the CUDA kernel
  get_fitnesses_neat             generate_illusion.py:478-673   -> sets genome.fitness for every genome
`neat_illusion` and the command line (generate_illusion.py:676-771) are NEAT's control plane and stay the reference's own:
its `eval_genomes` closure calls this module's `get_fitnesses_neat` instead of its own.
Differences, all listed in SURVEY.md "defects": no Colab import, no PNG hand-offs between stages (files are
written only for the best genome: best.png, best_black_bg.png, best_flow.png, enhanced.png), N=1 populations
work, Bands planes are reshaped to (h,w), colour uses
outputs 0..2 of 6-output configs, the dead 22nd PredNet forward is not computed.
"""
import os
import time

import numpy as np

from . import engine as engine_mod, genome as G, runtime
from ._lib import PAIR_POPULATION
from .grid import StructureType, create_grid, enhanced_image_grid  # noqa: F401  (re-exported, reference names)

REPEAT = 20  # generate_illusion.py:482
program_cache = G.ProgramCache()  # genome id -> flattened program; elites are re-submitted unchanged every generation


def _used_outputs(c_dim):
    return c_dim if c_dim > 1 else 1


_render_engines = {}


def get_image_from_cppn(inputs, genome, c_dim, w, h, config, bg=1, gradient=1, engine=None):
    """PIL image of one genome on the given grid planes (same signature as the reference + optional engine)."""
    from PIL import Image
    eng = engine
    if eng is None:    # a cached render-only context (tiny PredNet channels: only the CPPN kernel runs), one per image size
        key = (w, h, c_dim, "planes")
        if key not in _render_engines:
            _render_engines[key] = runtime.render_engine_factory(w, h, c_dim, 1)
        eng = _render_engines[key]
    eng.set_grid(grid=inputs)
    prog = G.flatten_genome_fast(genome, config, n_outputs=_used_outputs(c_dim))
    mode = engine_mod.render_mode_for(c_dim, gradient)
    img, _ = eng.render([prog], mode=mode, bg=float(bg))
    arr = img[0].cpu().numpy()
    if c_dim > 1:
        return Image.fromarray(arr)
    return Image.fromarray(arr[:, :, 0], "L")


def get_fitnesses_neat(structure, population, model_name, config, w, h, channels,
                       id=0, c_dim=3, best_dir=".", gradient=1, export_best=True, export_async=False):
    """population: [(genome_id, genome)].  On return every genome.fitness is a python float."""
    population = list(population)
    print("Calculating fitnesses of populations: ", len(population))
    eng = runtime.get_engine(w, h, channels, model_name, len(population))
    eng.set_grid(structure)
    mode = engine_mod.render_mode_for(c_dim, gradient)
    n_out = _used_outputs(c_dim)
    # each rank flattens its own shard, chunk by chunk, while the GPU evaluates the previous chunk; unchanged genomes
    # (elites) come out of the program cache
    t0 = time.perf_counter()
    fit = runtime.evaluate_genomes(eng, population, lambda gid, g: program_cache.get(gid, g, config, n_out),
                                   int(structure), mode, PAIR_POPULATION)
    dt = time.perf_counter() - t0
    program_cache.end_generation()
    print("evaluated %d genomes in %.1f ms (%.0f evals/s)" % (len(population), 1e3 * dt, len(population) / max(dt, 1e-9)))
    best_score, best_i = 0, 0
    for i, (_, genome) in enumerate(population):
        genome.fitness = float(fit[i])
        if genome.fitness >= best_score:  # generate_illusion.py:625 (NaN never wins, like the reference)
            best_score, best_i = genome.fitness, i
    print("scores", [[i, float(f)] for i, f in enumerate(fit)])
    if export_best and population and _is_export_rank():
        _export_best(eng, population[best_i], config, c_dim, gradient, best_dir, structure, export_async)
    print("best", best_score, best_i)
    return None


def _is_export_rank():
    """One process per GPU: every rank holds the full fitness vector after the all-gather, rank 0 writes the files."""
    import torch.distributed as dist
    return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0


ENHANCED_SIZE = 800  # generate_illusion.py:665-666


def _render_engine(w, h, c_dim, structure):
    """Render-only engine (tiny PredNet channels: only the CPPN kernel is used) for the enhanced mosaic; the mosaic's
    grid planes are uploaded once per structure."""
    key = (w, h, c_dim, int(structure))
    if key not in _render_engines:
        eng = runtime.render_engine_factory(w, h, c_dim, 1)
        eng.set_grid(grid=enhanced_image_grid(w, h, structure))
        _render_engines[key] = eng
    return _render_engines[key]


def _to_pil(arr, c_dim):
    from PIL import Image
    return Image.fromarray(arr) if c_dim > 1 else Image.fromarray(arr[:, :, 0], "L")


PNG_COMPRESS_LEVEL = 1   # zlib level of the exported PNGs: level 6 (PIL's default) costs ~0.2 s for the 800x800 mosaic
_export_pool = None
_export_pending = []


def wait_for_exports():
    """Block until every background file export (`export_async=True`) has been written."""
    while _export_pending:
        _export_pending.pop().result()


def _write_pngs(jobs):
    for image, path in jobs:
        image.save(path, "PNG", compress_level=PNG_COMPRESS_LEVEL)


def _export_best(eng, id_genome, config, c_dim, gradient, best_dir, structure=None, export_async=False):
    """The per-generation files of generate_illusion.py:650-671 for the best genome, without the PNG hand-offs:
    best.png, best_black_bg.png, best_flow.png (extension frame #1 with the flow vectors drawn,
    optical_flow.py:10-18,84-86) and enhanced.png (800x800 circle mosaic, lines 664-671; the grid is cached).
    The GPU work (three renders and one single-genome evaluation) takes a few milliseconds; PNG encoding is the
    expensive part and can run on a background thread (`export_async`, see wait_for_exports)."""
    global _export_pool
    from .optical_flow import draw_tracks
    os.makedirs(best_dir, exist_ok=True)
    prog = program_cache.get(id_genome[0], id_genome[1], config, _used_outputs(c_dim))
    mode = engine_mod.render_mode_for(c_dim, gradient)
    jobs = []
    for name, bg in (("best.png", 1.0), ("best_black_bg.png", 0.0)):
        img, _ = eng.render([prog], mode=mode, bg=bg)
        jobs.append((_to_pil(img[0].cpu().numpy(), c_dim), os.path.join(best_dir, name)))
    if structure is not None:
        eng.evaluate([prog], int(structure), mode, PAIR_POPULATION)      # one genome: frames + vectors of the winner
        dbg = eng.debug_buffers(1)
        vec = dbg["vectors"][0, :int(dbg["nvec"][0])]
        frame = dbg["frames"][1, 0]                                      # extension #1 = the image lucas_kanade draws on
        jobs.append((draw_tracks(_to_pil(frame, c_dim).convert("RGB"), vec), os.path.join(best_dir, "best_flow.png")))
        e_eng = _render_engine(ENHANCED_SIZE, ENHANCED_SIZE, c_dim, structure)
        img, _ = e_eng.render([prog], mode=mode, bg=1.0)
        jobs.append((_to_pil(img[0].cpu().numpy(), c_dim), os.path.join(best_dir, "enhanced.png")))
    if export_async:
        if _export_pool is None:
            import atexit
            from concurrent.futures import ThreadPoolExecutor
            _export_pool = ThreadPoolExecutor(max_workers=1)      # one writer: files of successive generations stay in order
            atexit.register(wait_for_exports)
        _export_pending.append(_export_pool.submit(_write_pngs, jobs))
    else:
        _write_pngs(jobs)
