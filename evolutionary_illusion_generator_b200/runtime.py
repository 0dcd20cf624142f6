"""Process-wide engine cache and genome-sharded evaluation (one process per GPU).

`evaluate_population` is what `get_fitnesses_neat` runs: flatten -> (shard) -> libeig -> (all-gather) -> floats.
Sharding follows SURVEY.md §8(e): rank r owns the contiguous block [r*ceil(N/G), (r+1)*ceil(N/G)) of the
population, in population order; PredNet weights and grid planes are replicated; one all-gather of the
padded per-rank fitness slices (NCCL over NVLink for CUDA tensors, gloo for the CPU-tensor unit tests) is the
only exchange.  Genomes are independent, so the result is bit-identical for every world size.
"""
import math
import os
import warnings

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, engine as engine_mod, genome as G

_engines = {}
engine_factory = engine_mod.Engine   # (w, h, channels, max_genomes) -> Engine; the CPU tests bind the host-compiled library here


def render_engine_factory(w, h, c_dim, max_genomes):
    """A render-only engine (eig_create_render: CPPN stage alone, any image size) from the same factory."""
    try:
        return engine_factory(w, h, (c_dim, 4, 4, 4), max_genomes, render_only=True)
    except TypeError:          # a test double without the keyword
        return engine_factory(w, h, (c_dim, 4, 4, 4), max_genomes)


def conv_policy():
    """EIG_CONV = auto | tc | simt (default auto): which convolution engine the drop-in entry points use.
    auto: the tcgen05 path; the exact-fp32 SIMT path when the library / device has no tensor-core path, and for the rest of
    the process once a weight file drives an activation out of the split-fp16 range (EIG_E_RANGE).  tc / simt: that
    path or an error."""
    p = os.environ.get("EIG_CONV", "auto").lower()
    if p not in ("auto", "tc", "simt"):
        raise ValueError("EIG_CONV must be auto, tc or simt (got %r)" % p)
    return p


def apply_conv_policy(eng):
    policy = conv_policy()
    if policy == "simt" or getattr(eng, "_range_fallback", False):
        eng.set_conv_mode(_lib.CONV_SIMT)
        eng.conv_mode = "simt"
        return eng
    try:
        eng.set_conv_mode(_lib.CONV_TC)
        eng.conv_mode = "tc"
    except _lib.EigError:
        if policy == "tc":
            raise
        eng.set_conv_mode(_lib.CONV_SIMT)
        eng.conv_mode = "simt"
    return eng


def with_range_fallback(eng, fn):
    """Run fn(); under EIG_CONV=auto an EIG_E_RANGE failure switches the engine to the exact-fp32 path and retries once."""
    try:
        return fn()
    except _lib.EigError as e:
        if e.code != _lib.EIG_E_RANGE or conv_policy() != "auto" or getattr(eng, "conv_mode", None) != "tc":
            raise
        warnings.warn("PredNet activations left the split-fp16 range of the tensor-core path; "
                      "this engine continues on the exact-fp32 SIMT convolution")
        eng._range_fallback = True
        apply_conv_policy(eng)
        return fn()


def get_engine(w, h, channels, model_name, max_genomes):
    """One Engine per (w, h, channels, weight file); re-created only when the population outgrows it.  The convolution
    engine follows `conv_policy()` and is re-applied whenever the engine is re-created."""
    key = (w, h, tuple(channels), model_name if isinstance(model_name, str) else id(model_name))
    eng = _engines.get(key)
    if eng is None or eng.max_genomes < max_genomes:
        grow, fell_back = 0, False
        if eng is not None:      # NEAT populations drift in size: grow geometrically instead of once per extra genome
            grow = eng.max_genomes + eng.max_genomes // 2
            fell_back = getattr(eng, "_range_fallback", False)
            eng.close()
        eng = engine_factory(w, h, channels, max(max_genomes, 8, grow))
        eng._range_fallback = fell_back
        apply_conv_policy(eng)
        eng.load_weights(model_name)
        _engines[key] = eng
    return eng


def shard_bounds(n, rank, world):
    """Contiguous block of ceil(n/world) genomes per rank (the last ranks may be short or empty)."""
    per = int(math.ceil(n / world)) if world > 0 else n
    lo = min(n, rank * per)
    return lo, min(n, lo + per), per


def gather_fitness(local, n_total, per, group=None):
    """All-gather the padded per-rank slices and cut the result back to population order.
    `local`: 1-D float64 tensor with this rank's fitness values (may be shorter than `per`)."""
    world = dist.get_world_size(group)
    pad = torch.full((per,), float("nan"), dtype=torch.float64, device=local.device)
    pad[:local.numel()] = local
    out = torch.empty((world * per,), dtype=torch.float64, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return out[:n_total]


def evaluate_population(eng, programs, structure, render_mode=engine_mod.RENDER_GRADIENT,
                        pair_mode=_lib.PAIR_POPULATION, group=None):
    """[FlatProgram] for the WHOLE population (same list on every rank) -> numpy float64 fitness vector."""
    n = len(programs)
    if n == 0:
        return np.zeros((0,), dtype=np.float64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return with_range_fallback(eng, lambda: eng.evaluate(programs, structure, render_mode, pair_mode))
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi, per = shard_bounds(n, rank, world)
    if hi > lo:
        local = torch.from_numpy(with_range_fallback(
            eng, lambda: eng.evaluate(programs[lo:hi], structure, render_mode, pair_mode))).to(eng.tdev)
    else:
        local = torch.zeros((0,), dtype=torch.float64, device=eng.tdev)
    return gather_fitness(local, n, per, group).cpu().numpy()


def evaluate_genomes(eng, population, flatten, structure, render_mode=engine_mod.RENDER_GRADIENT,
                     pair_mode=_lib.PAIR_POPULATION, group=None, chunk=None):
    """[(genome_id, genome)] for the WHOLE population (same list on every rank) -> numpy float64 fitness vector.
    Each rank flattens only its own shard (`flatten(genome_id, genome) -> FlatProgram`, e.g. `ProgramCache.get`), chunk
    by chunk, overlapped with the GPU evaluation of the previous chunk (`Engine.evaluate_streamed`); sharding and the
    single all-gather are those of `evaluate_population`."""
    population = list(population)
    n = len(population)
    if n == 0:
        return np.zeros((0,), dtype=np.float64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if (chunk is None and eng.stream_chunk(n) >= n) or (chunk is not None and chunk >= n):
            # one chunk: nothing to overlap - flatten, then the host entry point (one H2D, kernels, one D2H, one synchronisation)
            progs = [flatten(gid, g) for gid, g in population]
            return with_range_fallback(eng, lambda: eng.evaluate(progs, structure, render_mode, pair_mode))
        return with_range_fallback(
            eng, lambda: eng.evaluate_streamed(population, flatten, structure, render_mode, pair_mode, chunk)).cpu().numpy()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi, per = shard_bounds(n, rank, world)
    local = with_range_fallback(
        eng, lambda: eng.evaluate_streamed(population[lo:hi], flatten, structure, render_mode, pair_mode, chunk))
    return gather_fitness(local, n, per, group).cpu().numpy()
