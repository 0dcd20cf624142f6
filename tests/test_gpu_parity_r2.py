"""Round-2 parity tests on the B200 (`pytest -m gpu`), all through the C ABI: full bench populations against the
oracle, the folded up-sampled-R taps against the unfolded form, the precision profiles against the exact-fp32 path, the
largest reference-recorded case (512x512 Free), and the 2-rank NCCL path.  Nothing here reads /root/reference.

How fitness parity is stated (measured, profiles/r2/pass_ablation_c{2,3}_mixes512.md): fitness is a discontinuous
function of the uint8 frames (corner selection, LK status flags), so ANY two implementations whose frames differ in a
single LSB can disagree on a genome by more than 1e-3 - the exact-fp32 kernel with its nine taps summed in reverse order
already moves 3 of 512 C3 genomes beyond 1e-3 against itself.  The bars therefore are: frames within 1 LSB, the fraction
of differing bytes bounded, and fitness within 1e-3 relative for every genome that is not on an explicit, per-case
allow-list of genome ids (each entry a genome whose frames carry an LSB flip in that configuration)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT
from evolutionary_illusion_generator_b200 import _lib, engine as E, genome as G, weights as W

pytestmark = pytest.mark.gpu


def _progs(preset, c, idx):
    cfg = G.make_config(2, G.NEAT_PRESETS[preset]["num_outputs"])
    pop = [G.synthetic_genome(preset, i) for i in idx]
    return [G.flatten_genome(g, cfg, n_outputs=c if c > 1 else 1) for g in pop]


def _rel(got, want):
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-12)
    rel[(got == 0) & (want == 0)] = 0
    return np.where(np.isnan(got) & np.isnan(want), 0.0, rel)


def _frame_diff(a, b):
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    return int(d.max()), float((d > 0).mean())


# genomes of the full bench populations whose tensor-core frames carry an LSB flip that moves their fitness by more than
# 1e-3 against the CPU oracle (mode -> workload -> ids); measured on the B200, see the printed report of the test
FULL_POP_ALLOW = {
    "simt": {"c2": [], "c3": []},
    # c3 genome 29: one LSB flip in its extension frame changes the status of a tracked corner; 1.4e-2 (it is an outlier of the
    # exact 3-product path against the exact-fp32 GPU path too: profiles/r2/pass_ablation_c3_mixes512.md, ids 13 29 113 ...)
    "tc": {"c2": [], "c3": [29]},
}


@pytest.mark.parametrize("mode", ["simt", "tc"])
def test_full_bench_populations_vs_oracle(gpu_engine_factory, mode):
    """EVERY genome of the two 160x120 bench populations (BASELINE configs[1] pop 32, configs[2] pop 128) against the
    oracle's fitness (tests/golden/full_populations.npz, made by make_golden.py full): number of flow vectors equal,
    fitness within 1e-3 relative except the allow-listed genomes, first 8 genomes' frames within 1 LSB."""
    z = np.load(os.path.join(GOLDEN, "full_populations.npz"))
    for m in json.loads(str(z["meta"])):
        name, c, w, h, ch, n = m["name"], m["c_dim"], m["w"], m["h"], tuple(m["channels"]), m["n"]
        eng = gpu_engine_factory(w, h, ch, n)
        eng.set_conv_mode(_lib.CONV_TC if mode == "tc" else _lib.CONV_SIMT)
        eng.set_option("precision", 0)
        eng.set_grid(m["structure"])
        eng.load_weights(W.synthetic_predictor_weights(w, h, ch, seed=m["weight_seed"]))
        fit = eng.evaluate(_progs(m["preset"], c, range(n)), m["structure"])
        dbg = eng.debug_buffers(n)
        want = z["fitness_" + name]
        rel = _rel(fit, want)
        out = [int(i) for i in np.nonzero(rel > 1e-3)[0]]
        nvec_diff = [int(i) for i in np.nonzero(dbg["nvec"] != z["nvec_" + name])[0]]
        mx, frac = 0, 0.0
        for k in range(2):
            a = dbg["frames"][k, :8]
            a = a[..., 0] if c == 1 else a
            d = _frame_diff(a, z["frames_" + name][:, k])
            mx, frac = max(mx, d[0]), max(frac, d[1])
        print("full population %s [%s]: %d genomes, outside 1e-3: %s (allowed %s), nvec differs for %s, worst rel %.2e, "
              "median %.1e, frames of the first 8: max diff %d LSB, %.2e of bytes" %
              (name, mode, n, out, FULL_POP_ALLOW[mode][name], nvec_diff, rel.max(), np.median(rel), mx, frac))
        assert mx <= 1 and frac < 2e-4
        assert set(out) <= set(FULL_POP_ALLOW[mode][name]), (name, mode, out, rel[out])
        assert set(nvec_diff) <= set(FULL_POP_ALLOW[mode][name])
        assert (fit > 0).mean() > 0.4


@pytest.mark.parametrize("workload", ["c2", "c3"])
def test_folded_upsampled_taps_match_unfolded(gpu_engine_factory, workload):
    """ConvLSTM1/2 with the up-sampled-R taps folded to half resolution (tap-masked Z convolution + epilogue add) against
    the same convolutions over the 2x2-replicated R: the same products summed in another order - frames within 1 LSB on a
    handful of bytes, fitness equal within the fp32-reordering floor; both against the exact-fp32 path."""
    preset, c, ch = ("circles_bw", 1, (1, 16, 32, 64)) if workload == "c2" else ("circles", 3, (3, 48, 96, 192))
    w, h, n = 160, 120, 16
    eng = gpu_engine_factory(w, h, ch, n)
    eng.set_grid(1)
    eng.load_weights(W.synthetic_predictor_weights(w, h, ch, seed=0))
    progs = _progs(preset, c, range(40, 40 + n))
    eng.set_conv_mode(_lib.CONV_SIMT)
    f_ref = eng.evaluate(progs, 1)
    fr_ref = eng.debug_buffers(n)["frames"][:2].copy()
    eng.set_conv_mode(_lib.CONV_TC)
    eng.set_option("precision", 0)
    res = {}
    for fold in (0, 1):
        eng.set_option("fold", fold)
        l0 = eng.lib.eig_launch_count()
        f = eng.evaluate(progs, 1)
        launches = eng.lib.eig_launch_count() - l0
        assert np.array_equal(f, eng.evaluate(progs, 1))                       # deterministic (second call replays the graph)
        res[fold] = (f, eng.debug_buffers(n)["frames"][:2].copy(), launches)
    eng.set_option("fold", -1)
    assert res[1][2] == res[0][2] + 2 * 21                                     # two Z launches per PredNet step
    for fold in (0, 1):
        mx, frac = _frame_diff(res[fold][1], fr_ref)
        rel = _rel(res[fold][0], f_ref)
        print("fold %d %s: frames vs fp32 path max %d LSB, %.2e of bytes; fitness rel err max %.2e, outside 1e-3: %d of %d" %
              (fold, workload, mx, frac, rel.max(), int((rel > 1e-3).sum()), n))
        assert mx <= 1 and frac < 1e-4
        assert (rel > 1e-3).sum() <= 1
    mx, frac = _frame_diff(res[1][1], res[0][1])
    assert mx <= 1 and frac < 1e-4


@pytest.mark.parametrize("workload", ["c2", "c3"])
def test_zero_state_skip_is_bit_identical(gpu_engine_factory, workload):
    """Step 0 of a sequence skips the K blocks that only hold the zero state (E- with P = 0, h = 0): exact zeros, so frames
    and fitness must not change by a single bit, folded or not."""
    preset, c, ch = ("circles_bw", 1, (1, 16, 32, 64)) if workload == "c2" else ("circles", 3, (3, 48, 96, 192))
    w, h, n = 160, 120, 8
    eng = gpu_engine_factory(w, h, ch, n)
    eng.set_conv_mode(_lib.CONV_TC)
    eng.set_option("precision", 0)
    eng.set_grid(1)
    eng.load_weights(W.synthetic_predictor_weights(w, h, ch, seed=0))
    progs = _progs(preset, c, range(70, 70 + n))
    for fold in (0, 1):
        eng.set_option("fold", fold)
        out = {}
        for skip in (0, 1):
            eng.set_option("skip_zero_state", skip)
            f = eng.evaluate(progs, 1)
            out[skip] = (f, eng.debug_buffers(n)["frames"].copy())
        assert np.array_equal(out[0][0], out[1][0], equal_nan=True), (workload, fold)
        assert np.array_equal(out[0][1], out[1][1]), (workload, fold)
    eng.set_option("fold", -1)
    eng.set_option("skip_zero_state", 1)


def test_precision_profiles_stay_within_one_lsb(gpu_engine_factory):
    """eig_set_option("precision"): 0 exact (3 products everywhere), 1 balanced (single product in layers 2 and 3),
    2 fast (single product everywhere).  Against the exact-fp32 path on 32 C3 genomes: frames never differ by more than
    1 LSB, the differing fraction grows in the measured order (profiles/r2/pass_ablation_c3_mixes512.md), and the bulk of
    the population stays within 1e-3."""
    w, h, ch, n = 160, 120, (3, 48, 96, 192), 32
    eng = gpu_engine_factory(w, h, ch, n)
    eng.set_grid(1)
    eng.load_weights(W.synthetic_predictor_weights(w, h, ch, seed=0))
    progs = _progs("circles", 3, range(200, 200 + n))
    eng.set_conv_mode(_lib.CONV_SIMT)
    f_ref = eng.evaluate(progs, 1)
    fr_ref = eng.debug_buffers(n)["frames"][:2].copy()
    eng.set_conv_mode(_lib.CONV_TC)
    bars = {0: (1e-4, 2), 1: (3e-4, 4), 2: (1.5e-3, 8)}    # (fraction of differing bytes, genomes outside 1e-3) per profile
    for prof in (0, 1, 2):
        eng.set_option("precision", prof)
        f = eng.evaluate(progs, 1)
        mx, frac = _frame_diff(eng.debug_buffers(n)["frames"][:2], fr_ref)
        rel = _rel(f, f_ref)
        print("precision %d: frames max %d LSB, %.2e of bytes differ; fitness rel err median %.1e max %.1e, outside 1e-3: %d of %d"
              % (prof, mx, frac, np.median(rel), rel.max(), int((rel > 1e-3).sum()), n))
        assert mx <= 1 and frac < bars[prof][0] and (rel > 1e-3).sum() <= bars[prof][1]
    eng.set_option("precision", 0)
    with pytest.raises(_lib.EigError):
        eng.set_option("precision", 3)


@pytest.mark.parametrize("mode", ["simt", "tc"])
def test_largest_reference_recorded_case(gpu_engine_factory, mode):
    """BASELINE configs[4] shape - 512x512 colour, Free scoring branch, three LK pyramid levels - against the fitness and
    the frames the REFERENCE's own get_fitnesses_neat produced (tests/golden/reference_pipeline_large.npz, recorded under
    the Chainer stand-in by make_golden.py reference-large)."""
    path = os.path.join(GOLDEN, "reference_pipeline_large.npz")
    if not os.path.isfile(path):
        pytest.skip("reference_pipeline_large.npz not generated")
    z = np.load(path)
    for m in json.loads(str(z["meta"])):
        w, h, ch, c, n = m["w"], m["h"], tuple(m["channels"]), m["c_dim"], m["n"]
        cfg = G.make_config(2, G.NEAT_PRESETS[m["preset"]]["num_outputs"])
        pop = G.synthetic_population(m["preset"], n, evolved=m["evolved"])
        progs = [G.flatten_genome(g, cfg, n_outputs=c) for _, g in pop]
        eng = gpu_engine_factory(w, h, ch, n)
        eng.set_conv_mode(_lib.CONV_TC if mode == "tc" else _lib.CONV_SIMT)
        eng.set_option("precision", 0)
        eng.set_grid(m["structure"])
        eng.load_weights(W.synthetic_predictor_weights(w, h, ch, seed=m["weight_seed"]))
        fit = eng.evaluate(progs, m["structure"], E.render_mode_for(c, m["gradient"]))
        ref = z["fitness_" + m["name"]]
        frames = eng.debug_buffers(n)["frames"]
        for k in range(2):
            mx, frac = _frame_diff(frames[k], z["frames_" + m["name"]][:, k])
            assert mx <= 1 and frac < 1e-3, (m["name"], k, mx, frac)
        print("%s [%s]: gpu %s reference %s" % (m["name"], mode, fit, ref))
        assert np.allclose(fit, ref, rtol=1e-3, atol=1e-9, equal_nan=True), (m["name"], mode, fit, ref)


_NCCL_WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["EIG_ROOT"])
from evolutionary_illusion_generator_b200 import _lib, engine as E, genome as G, runtime, weights as W
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
w, h, ch, n = 160, 120, (1, 16, 32, 64), 11           # ragged: 6 + 5 genomes
eng = E.Engine(w, h, ch, n, device=rank)
eng.set_conv_mode(_lib.CONV_TC)
eng.set_grid(1)
eng.load_weights(W.synthetic_predictor_weights(w, h, ch, seed=0))
cfg = G.make_config(2, 1)
pop = [(i, G.synthetic_genome("circles_bw", i)) for i in range(n)]
progs = [G.flatten_genome(g, cfg, n_outputs=1) for _, g in pop]
sharded = runtime.evaluate_population(eng, progs, 1)               # each rank its shard + one NCCL all-gather
cache = G.ProgramCache()
streamed = runtime.evaluate_genomes(eng, pop, lambda gid, g: cache.get(gid, g, cfg, 1), 1)
single = eng.evaluate(progs, 1)                                    # the whole population on this rank's GPU alone
np.save(os.path.join(os.environ["EIG_OUT"], "r%d.npy" % rank), np.stack([sharded, streamed, single]))
dist.barrier()
dist.destroy_process_group()
eng.close()
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_rank_nccl_gather_equals_single_gpu(tmp_path):
    """SURVEY.md §4 item iv on the real transport: two processes, one GPU each, NCCL.  The vector every rank holds after
    the all-gather equals, bit for bit, the evaluation of the whole population on one GPU - on both ranks, for the
    up-front and the streamed route."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_NCCL_WORKER)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), EIG_ROOT=ROOT, EIG_OUT=str(tmp_path))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(r0, r1, equal_nan=True)
    assert np.array_equal(r0[0], r0[2], equal_nan=True) and np.array_equal(r0[1], r0[2], equal_nan=True)
    assert np.any(r0[2] > 0)
