"""Genome sharding + all-gather with world_size 2 on the gloo backend (CPU).  Each rank evaluates its shard
with the emulator-built kernel sources; the gathered vector must equal the single-process result bit for bit
(genomes are independent)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import EMU_SO


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from evolutionary_illusion_generator_b200 import _lib, engine as E, genome as G, runtime, weights as W
    torch.set_num_threads(1)
    w, h, ch = 64, 64, (1, 4, 8, 8)
    eng = E.Engine(w, h, ch, n, lib=_lib.EigLibrary(EMU_SO))
    eng.set_grid(1)
    eng.load_weights(W.synthetic_weights(w, h, ch, seed=2))
    cfg = G.make_config(2, 1)
    progs = [G.flatten_genome(G.synthetic_genome("circles_bw", i), cfg, n_outputs=1) for i in range(n)]
    fit = runtime.evaluate_population(eng, progs, 2)
    lo, hi, per = runtime.shard_bounds(n, rank, world)
    np.save(os.path.join(out_dir, "fit_%d.npy" % rank), fit)
    # the streamed route get_fitnesses_neat takes: every rank flattens only its shard, one genome per chunk here
    cache = G.ProgramCache()
    pop = [(i, G.synthetic_genome("circles_bw", i)) for i in range(n)]
    streamed = runtime.evaluate_genomes(eng, pop, lambda gid, g: cache.get(gid, g, cfg, 1), 2, chunk=1)
    assert cache.misses + cache.fast == hi - lo     # this rank flattened its own shard only
    np.save(os.path.join(out_dir, "fit_streamed_%d.npy" % rank), streamed)
    np.save(os.path.join(out_dir, "shard_%d.npy" % rank), np.array([lo, hi, per]))
    # the whole drop-in call under torch.distributed: every rank ends with the same genome.fitness, rank 0 alone writes files
    from evolutionary_illusion_generator_b200 import generate_illusion as GI
    emu = _lib.EigLibrary(EMU_SO)
    runtime.engine_factory = lambda w_, h_, ch_, n_, **kw: E.Engine(w_, h_, ch_, n_, lib=emu, **kw)
    GI.ENHANCED_SIZE = 64
    model = os.path.join(out_dir, "model_%d.npz" % rank)
    W.save_npz(model, W.synthetic_weights(w, h, ch, seed=2))
    pop2 = [(i, G.synthetic_genome("circles_bw", i)) for i in range(n)]
    GI.get_fitnesses_neat(GI.StructureType.Free, pop2, model, cfg, w, h, ch, c_dim=1,
                          best_dir=os.path.join(out_dir, "best_%d" % rank))
    np.save(os.path.join(out_dir, "fit_neat_%d.npy" % rank), np.array([g.fitness for _, g in pop2]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_evaluation_equals_single_process(emu_lib, tmp_path):
    n, world = 3, 2   # ragged: rank 0 gets 2 genomes, rank 1 gets 1
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    from evolutionary_illusion_generator_b200 import engine as E, genome as G, weights as W
    w, h, ch = 64, 64, (1, 4, 8, 8)
    eng = E.Engine(w, h, ch, n, lib=emu_lib)
    eng.set_grid(1)
    eng.load_weights(W.synthetic_weights(w, h, ch, seed=2))
    cfg = G.make_config(2, 1)
    progs = [G.flatten_genome(G.synthetic_genome("circles_bw", i), cfg, n_outputs=1) for i in range(n)]
    single = eng.evaluate(progs, 2)
    f0, f1 = np.load(tmp_path / "fit_0.npy"), np.load(tmp_path / "fit_1.npy")
    assert np.array_equal(f0, f1) and f0.shape == (n,)
    assert np.array_equal(f0, single) and np.any(single > 0)
    assert np.array_equal(np.load(tmp_path / "fit_streamed_0.npy"), single)
    assert np.array_equal(np.load(tmp_path / "fit_streamed_1.npy"), single)
    pop = [(i, G.synthetic_genome("circles_bw", i)) for i in range(n)]
    for chunk in (None, 2):
        one = eng.evaluate_streamed(pop, lambda gid, g: G.flatten_genome(g, cfg, n_outputs=1), 2, chunk=chunk)
        assert np.array_equal(one.numpy(), single)
    eng.set_grid(2)                                            # get_fitnesses_neat(Free): Free grid and Free scoring
    single_free = eng.evaluate(progs, 2)
    assert np.array_equal(np.load(tmp_path / "fit_neat_0.npy"), single_free)
    assert np.array_equal(np.load(tmp_path / "fit_neat_1.npy"), single_free)
    assert sorted(os.listdir(tmp_path / "best_0")) == ["best.png", "best_black_bg.png", "best_flow.png", "enhanced.png"]
    assert not os.path.exists(tmp_path / "best_1")
    assert list(np.load(tmp_path / "shard_0.npy")) == [0, 2, 2] and list(np.load(tmp_path / "shard_1.npy")) == [2, 3, 2]


def test_gather_pads_empty_ranks(tmp_path):
    mp.spawn(_gather_worker, args=(3, _free_port(), str(tmp_path)), nprocs=3, join=True)
    for r in range(3):
        assert list(np.load(tmp_path / ("g_%d.npy" % r))) == [10.0, 11.0]


def _gather_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from evolutionary_illusion_generator_b200 import runtime
    n = 2                                    # 2 genomes on 3 ranks: the last rank owns nothing
    lo, hi, per = runtime.shard_bounds(n, rank, world)
    local = torch.tensor([10.0 + i for i in range(lo, hi)], dtype=torch.float64)
    out = runtime.gather_fitness(local, n, per)
    np.save(os.path.join(out_dir, "g_%d.npy" % rank), out.numpy())
    dist.destroy_process_group()
