#!/usr/bin/env python
"""bench.py - genome fitness evaluations per second on N B200s (BASELINE.json metric).

One "step" = one pass of the hot path (CPPN render -> PredNet 20+1 steps -> Shi-Tomasi/LK flow -> score) over one
synthetic population shard per GPU, with one all-gather (NCCL) of the fitness slices inside every step when N > 1.

The contract line is BASELINE.json configs[2] ("c3": pop 128, neat_configs/circles.txt colour, 160x120, PredNet
channels 3,48,96,192 = the reference's `--channels` default, /root/reference/generate_illusion.py:734) - the config the
metric's "1/2/4/8 B200" sweep is quoted on - weak scaling with 128 genomes per GPU.  The same JSON line carries, under
"also", one sub-record per further configuration, each measured the same way on the same box:
  c2           BASELINE configs[1]: pop 32 gray 160x120 (channels 1,16,32,64), per GPU
  c3_strong    configs[2] split over the ranks (pop 128 / N per GPU); N > 1 only
  c4, c5       configs[3] (pop 256, bands, 320x240) and configs[4] (pop 1024, free, 512x512) - their "8xB200" shape,
               i.e. pop / 8 genomes per GPU; measured when N = 8 (or with --also c4,c5)
  c3_precision1, c3_precision2   N = 1: the contract workload under the two OPT-IN precision profiles (fewer tensor-core
               products per MAC; frames stay within 1 LSB, more bytes differ - profiles/r2/pass_ablation_*.md).  The
               contract line itself always runs the exact default (three products per MAC).

  value    : evals/s with the flattened genomes already resident in HBM; CUDA events per step on the launching stream
             (L2 flushed between steps, untimed), max over ranks.
  e2e      : the same through the host entry point `eig_eval_host` (pinned host genome blob -> H2D -> kernels -> D2H
             fitness), wall clock with a device sync, max over ranks.  `e2e_from_genomes` (rank 0, N = 1) starts one
             level higher, at the genome objects `get_fitnesses_neat` receives (flatten + pack + eig_eval_host).
  parity   : BEFORE anything is timed, the first k genomes of rank 0's shard are evaluated by the CPU oracle and
             compared with the GPU fitness (1e-3 relative, the `north_star` tolerance).  A failing gate aborts the run.
  gathered_check : N > 1: rank 0 re-evaluates the WHOLE global population on its own GPU and compares the all-gathered
             vector bit for bit.
  roofline : per-kernel-class CUDA-event times of an instrumented pass of the same step (eig_profile_*).
  cpu_baseline : the oracle (port of the reference's path; Chainer is not installable) on a bounded sample.

`--impl reference` times the reference's CPU path (oracle port, all host threads) on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU oracle of the parity gate (rank 0) would then take minutes per
# genome.  Give each rank its share of the host cores instead - before numpy / torch load their thread pools.
if os.environ.get("OMP_NUM_THREADS") == "1" and os.environ.get("LOCAL_WORLD_SIZE"):
    _share = max(1, (os.cpu_count() or 1) // max(1, int(os.environ["LOCAL_WORLD_SIZE"])))
    if os.environ.get("RANK", "0") == "0":
        _share = max(_share, min(os.cpu_count() or 1, 16))
    os.environ["OMP_NUM_THREADS"] = str(_share)
    os.environ["MKL_NUM_THREADS"] = str(_share)

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (preset, c_dim, channels, w, h, structure, pop per GPU, MACs per layer-0 pixel per PredNet step)
    "c2": ("circles_bw", 1, (1, 16, 32, 64), 160, 120, 1, 32, 37269),
    "c3": ("circles", 3, (3, 48, 96, 192), 160, 120, 1, 128, 335421),
    # per-GPU shares of BASELINE configs[3] / configs[4] on 8 GPUs (pop 256 -> 32, pop 1024 -> 128): stress / roofline
    "c4": ("bands", 3, (3, 48, 96, 192), 320, 240, 0, 32, 335421),
    "c5": ("free", 3, (3, 48, 96, 192), 512, 512, 2, 128, 335421),
}
BASELINE_CONFIG = {"c2": "configs[1]", "c3": "configs[2]", "c4": "configs[3] (pop 256 / 8 GPUs)", "c5": "configs[4] (pop 1024 / 8 GPUs)"}
STRUCTURE_NAMES = {0: "Bands", 1: "Circles", 2: "Free", 3: "CirclesFree"}
MIN_TIMED_S = 1.0   # every point is timed for at least this long (steps are raised internally for the short workloads)


def tc_kernel_macs(ch):
    """Algorithmic MACs per layer-0 pixel per PredNet step that the tcgen05 kernel covers (SURVEY.md §8 a-7 counts):
    everything except ConvA1, ConvP0 and the E0/h0 taps of ConvLSTM0, which run on the layer-0 SIMT kernels.
    The up-sampled-R taps of every ConvLSTM are counted at their reference cost (9 taps at full resolution), however the
    kernel evaluates them (folded to half resolution)."""
    c0, c1, c2, c3 = ch
    total = 0.0
    total += 9 * 2 * c1 * c2 / 4.0 + 9 * 2 * c2 * c3 / 16.0                      # ConvA2, ConvA3
    total += 9 * (2 * c1 + c2 + c1) * 4 * c1 / 4.0                                # ConvLSTM1
    total += 9 * (2 * c2 + c3 + c2) * 4 * c2 / 16.0 + 9 * (2 * c3 + c3) * 4 * c3 / 64.0   # ConvLSTM2, ConvLSTM3
    total += 9 * c1 * c1 / 4.0 + 9 * c2 * c2 / 16.0 + 9 * c3 * c3 / 64.0          # ConvP1..3
    total += 9 * c1 * 4 * c0                                                      # R1 taps of ConvLSTM0
    if c1 >= 32:
        total += 9 * 2 * c0 * c1                                                  # ConvA1 (on tcgen05 for wide first layers)
    return total


def tc_dead_macs(ch):
    """ConvP2 / ConvP3 of the final step feed nothing (no next step): not launched, not counted."""
    return 9 * ch[2] * ch[2] / 16.0 + 9 * ch[3] * ch[3] / 64.0


USEFUL_STEPS = 21  # 20 static frames + 1 self-fed (the reference's 22nd forward is never read)
METRIC = "NEAT genome fitness evals/sec (CPPN+PredNet+flow) @160x120"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf=d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)),
                    source="measured")
    return dict(hbm=6650.0, tf=1590.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def build_population(preset, c_dim, n, start):
    from evolutionary_illusion_generator_b200 import genome as G
    n_out = G.NEAT_PRESETS[preset]["num_outputs"]
    cfg = G.make_config(2, n_out)
    pop = [G.synthetic_genome(preset, start + i) for i in range(n)]
    progs = [G.flatten_genome(g, cfg, n_outputs=c_dim if c_dim > 1 else 1) for g in pop]
    return cfg, pop, progs


_weights_cache = {}


def workload_weights(workload):
    from evolutionary_illusion_generator_b200 import weights as W
    preset, c_dim, ch, w, h = WORKLOADS[workload][:5]
    key = (w, h, ch)
    if key not in _weights_cache:
        _weights_cache.clear()          # one set at a time: the 512x512 peephole maps are large
        _weights_cache[key] = W.synthetic_predictor_weights(w, h, ch, seed=0)
    return _weights_cache[key]


def oracle_fitness(workload, genomes_from, n, threads, flow_impl="cv2"):
    """The reference's CPU path (oracle port) on `n` genomes of the workload -> (fitness vector, seconds)."""
    import torch
    from oracle import pipeline as OPL
    preset, c_dim, ch, w, h, structure, _, _ = WORKLOADS[workload]
    torch.set_num_threads(threads)
    wts = workload_weights(workload)
    cfg, pop, _ = build_population(preset, c_dim, n, genomes_from)
    gc = cfg.genome_config
    t0 = time.perf_counter()
    fit = OPL.evaluate_population(pop, gc.input_keys, gc.output_keys, structure, wts, w, h, ch, c_dim, flow_impl=flow_impl)
    return np.asarray(fit, dtype=np.float64), time.perf_counter() - t0


def oracle_evals_per_s(workload, n_sample, threads, flow_impl="cv2"):
    _, dt = oracle_fitness(workload, 0, n_sample, threads, flow_impl)
    return n_sample / dt, dt


def default_ref_sample(workload):
    return {"c2": 8, "c3": 2, "c4": 1, "c5": 1}[workload]


def workload_config(workload, pop, world, scaling, conv):
    preset, c_dim, ch, w, h, structure, _, _ = WORKLOADS[workload]
    return {"workload": workload, "baseline_config": BASELINE_CONFIG[workload], "pop_per_gpu": pop, "global_pop": world * pop,
            "resolution": "%dx%d" % (w, h), "channels": list(ch), "neat_config": preset,
            "structure": STRUCTURE_NAMES[structure], "prednet_steps": USEFUL_STEPS, "conv": conv,
            "weights": "synthetic_predictor_weights seed 0 (LeCun-normal, layer 0 shaped as an error integrator)",
            "l2": "flushed between steps (256 MiB memset, untimed)",
            "parallelism": "genome-sharded dp%d + 1 all-gather/step" % world, "scaling": scaling}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    preset, c_dim, ch, w, h, structure, pop, _ = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    n_sample = args.ref_sample or default_ref_sample(args.workload)
    oracle_evals_per_s(args.workload, 1, cores)          # warm-up (thread pools, grids)
    secs = 0.0
    for _ in range(args.steps):
        _, dt = oracle_evals_per_s(args.workload, n_sample, cores)
        secs += dt
    value = args.steps * n_sample / secs
    sample = ("%d genomes per step of workload %s (render + 22 PredNet forwards + cv2 LK + scoring, in memory; genomes are "
              "evaluated one after the other exactly as the reference does, so evals/s does not depend on the sample size), "
              "torch-CPU fp32 oracle port of the Chainer path, %d threads" % (n_sample, args.workload, cores))
    cfg = workload_config(args.workload, pop, args.gpus, "weak", "cpu")
    cfg["sample_genomes_per_step"] = n_sample
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


class Dist:
    def __init__(self):
        import torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        import torch
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def close(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()


def measure(D, workload, pop, steps, warmup, scaling, conv, flush, main, parity_ref, cpu_sample, from_genomes, precision=None):
    """One configuration, measured on every rank; rank 0 returns the record (other ranks None)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from evolutionary_illusion_generator_b200 import _lib, engine as E, genome as G
    preset, c_dim, ch, w, h, structure, _, macs = WORKLOADS[workload]
    rank, world, dev = D.rank, D.world, D.dev
    eng = E.Engine(w, h, ch, pop, device=D.local)
    eng.set_conv_mode(_lib.CONV_TC if conv == "tc" else _lib.CONV_SIMT)
    if precision is not None:
        eng.set_option("precision", precision)
    eng.set_grid(structure)
    eng.load_weights(workload_weights(workload))
    cfg, genomes, progs = build_population(preset, c_dim, pop, rank * pop)
    blob, offsets, max_slots = G.pack_population(progs)
    resident = eng.upload_programs(progs)
    fit_dev = torch.empty((pop,), dtype=torch.float64, device=dev)
    gathered = torch.empty((world * pop,), dtype=torch.float64, device=dev)
    pin_blob = torch.from_numpy(blob).pin_memory()
    pin_off = torch.from_numpy(offsets).pin_memory()
    pin_fit = torch.empty((pop,), dtype=torch.float64).pin_memory()
    blob_np, off_np, fit_np = pin_blob.numpy(), pin_off.numpy(), pin_fit.numpy()

    def step_resident():
        eng.evaluate_resident(resident, structure, out=fit_dev)
        if world > 1:
            dist.all_gather_into_tensor(gathered, fit_dev)

    def step_host():
        eng.evaluate_host(blob_np, off_np, max_slots, structure, out=fit_np)
        if world > 1:
            fit_dev.copy_(pin_fit, non_blocking=True)
            dist.all_gather_into_tensor(gathered, fit_dev)
            torch.cuda.synchronize()

    # ---- parity gate, before anything is timed (rank 0)
    parity = None
    if parity_ref is not None:
        ok = True
        if rank == 0:
            want, dt = parity_ref
            k = min(len(want), pop)
            want = want[:k]
            eng.evaluate_resident(resident, structure, out=fit_dev)
            torch.cuda.synchronize()
            got = fit_dev[:k].cpu().numpy()
            both_nan = np.isnan(got) & np.isnan(want)       # a zero-length flow vector is NaN in the reference too
            rel = np.where(both_nan, 0.0, np.abs(got - want) / np.maximum(np.abs(want), 1e-9))
            within = int(np.sum(both_nan | (rel <= 1e-3) | (np.abs(got - want) <= 1e-9)))
            parity = {"checked": k, "within_1e-3": within, "ok": within == k, "max_rel_err": float(np.nanmax(rel)),
                      "against": "CPU oracle (oracle/pipeline.py, torch-CPU fp32 PredNet port + cv2 LK), same genomes and weights, %.1f s" % dt,
                      "nonzero": int((want > 0).sum())}
            ok = within == k
        if world > 1:
            flag = torch.tensor([1 if ok else 0], device=dev)
            dist.broadcast(flag, 0)
            ok = bool(flag.item())
        if not ok and main:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "error": "parity gate failed before timing", "parity": parity,
                                  "config": workload_config(workload, pop, world, scaling, conv)}))
            eng.close()
            D.close()
            sys.exit(1)

    for _ in range(max(warmup, 3)):
        step_resident()
        flush.zero_()
    D.barrier()
    # raise the step count so that the timed region lasts >= MIN_TIMED_S (the contract line keeps the driver's K)
    if not main:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step_resident(); b.record(); torch.cuda.synchronize()
        est = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(est, op=dist.ReduceOp.MAX)
        steps = max(steps, int(np.ceil(1e3 * MIN_TIMED_S / max(float(est[0]), 1e-3))))
    sampler = ClockSampler(D.local)
    if rank == 0:
        sampler.start()
    launches0 = eng.lib.eig_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    D.barrier()
    for a, b in ev:
        a.record()
        step_resident()
        b.record()
        flush.zero_()          # L2 flush between timed steps (not inside any event pair)
    D.barrier()
    launches = eng.lib.eig_launch_count() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    # end-to-end through the host entry point
    e2e_steps = steps
    for _ in range(2):
        step_host()
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    D.barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    tmax = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(tmax[0]), float(tmax[1])
    step_resident()
    torch.cuda.synchronize()
    fit_host = fit_dev.cpu().numpy()

    # ---- N > 1: the gathered vector against a single-GPU evaluation of the same global population (rank 0)
    gathered_check = None
    if world > 1:
        all_host = gathered.cpu().numpy()
        if rank == 0:
            single = np.empty_like(all_host)
            for r in range(world):
                _, _, pr = build_population(preset, c_dim, pop, r * pop)
                single[r * pop:(r + 1) * pop] = eng.evaluate(pr, structure)
            same = np.array_equal(single, all_host, equal_nan=True)
            gathered_check = {"genomes": int(all_host.size), "bit_equal_to_single_gpu": bool(same),
                              "max_abs_diff": float(np.nanmax(np.abs(single - all_host))) if not same else 0.0}
        D.barrier()

    # ---- genome objects -> fitness (what get_fitnesses_neat does per generation), rank 0, N = 1
    e2e_genomes = None
    if from_genomes and rank == 0 and world == 1:
        e2e_genomes = time_from_genomes(eng, genomes, cfg, c_dim, structure, steps)

    # ---- instrumented pass: per-kernel-class device time of the same step (rank 0 reports)
    ms = (C.c_double * 8)()
    cnt = (C.c_int64 * 8)()
    prof_steps = 2
    eng.lib.check(eng.lib.eig_profile_begin(eng.ctx))
    for _ in range(prof_steps):
        eng.evaluate_resident(resident, structure, out=fit_dev)
    eng.lib.check(eng.lib.eig_profile_end(eng.ctx, ms, cnt))
    cls_names = ["render", "conv_simt", "conv_tcgen05", "elementwise", "flow", "score", "layer0_fused"]
    cls_ms = {cls_names[i]: ms[i] / prof_steps for i in range(7)}
    cls_n = {cls_names[i]: int(cnt[i] // prof_steps) for i in range(7)}
    eng.close()
    if rank != 0:
        return None

    peaks = measured_peaks()
    total = world * pop * steps
    value = total / (dev_ms / 1e3)
    e2e = world * pop * e2e_steps / (e2e_ms / 1e3)
    flop_step = 2.0 * (tc_kernel_macs(ch) * USEFUL_STEPS - tc_dead_macs(ch)) * w * h * pop   # algorithmic FLOP of the tcgen05 launches, one GPU
    conv_ms = cls_ms["conv_simt"] + cls_ms["conv_tcgen05"]
    conv_n = cls_n["conv_simt"] + cls_n["conv_tcgen05"]
    achieved = flop_step / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    traffic, tsrc = None, None
    for rel in ("profiles/r2/tc_traffic.json", "profiles/r1/tc_traffic.json"):
        tpath = os.path.join(ROOT, rel)
        if conv == "tc" and os.path.isfile(tpath) and pop == WORKLOADS[workload][6]:
            t = json.load(open(tpath)).get(workload, {}).get("mean_dram_bytes_per_launch")
            if t:
                traffic, tsrc = t, "ncu --set full, %s (mean over the conv launches of one PredNet step)" % rel
                break
    roofline = {"bound": "tensor", "kernel": "conv3x3_tc_kernel (%s)" % ("tcgen05 cta_group::2, 3-pass split fp16" if conv == "tc" else "fp32 SIMT"),
                "achieved": achieved, "peak": peaks["tf"], "unit": "TFLOP/s", "frac": achieved / peaks["tf"],
                "traffic": traffic, "traffic_source": tsrc,
                "peak_source": peaks["source"] + " bf16 dense (sustained)",
                "flop_per_launch": flop_step / max(conv_n, 1), "launches_per_step": conv_n,
                "avg_launch_us": 1e3 * conv_ms / max(conv_n, 1),
                "mma_passes": 3 if conv == "tc" else None,
                "frac_of_3pass_ceiling": 3.0 * achieved / peaks["tf"] if conv == "tc" else None,
                "class_ms_per_step": cls_ms, "class_launches_per_step": cls_n,
                "whole_path_gflop_per_genome": 2.0 * macs * w * h * USEFUL_STEPS / 1e9,
                "whole_path_tflops": 2.0 * macs * w * h * USEFUL_STEPS * world * pop * steps / (dev_ms / 1e3) / 1e12,
                "timing": "per-launch CUDA events of an instrumented pass of the same step (library launch instrumentation; "
                          "graph replay, side-stream overlap and programmatic dependent launch are off in that pass, "
                          "so the class times are upper bounds of their share of ms_per_step)",
                "note": "achieved counts each algorithmic MAC once (up-sampled taps at their reference cost); fp32-grade accuracy "
                        "needs 3 fp16 MMAs per MAC (hi*hi + hi*lo + lo*hi), so 1/3 of the dense 16-bit peak is the ceiling of this kernel"}
    cpu = None
    if cpu_sample > 0 and world == 1:
        cores = os.cpu_count() or 1
        oracle_evals_per_s(workload, 1, cores)
        v, dt = oracle_evals_per_s(workload, cpu_sample, cores)
        cpu = {"value": v, "unit": "evals/s", "cores": cores, "kind": "port",
               "sample": "%d genomes of workload %s through the oracle (torch-CPU fp32 PredNet port, cv2 LK), "
                         "%.1f s" % (cpu_sample, workload, dt)}
    rec = {"metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": steps,
           "warmup": max(warmup, 3), "ms_per_step": dev_ms / steps, "higher_is_better": True,
           "scaling": scaling, "vs_baseline": None,
           "dtype": "fp32 (3-pass split-fp16 tcgen05, fp32 accumulate)" if conv == "tc" else "fp32",
           "data": "synthetic", "config": workload_config(workload, pop, world, scaling, conv),
           "e2e": {"value": e2e, "unit": "evals/s", "h2d_bytes_per_step": int(blob.nbytes + offsets.nbytes),
                   "d2h_bytes_per_step": int(8 * pop), "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps},
           "gpu_launches": int(launches), "clocks": clocks, "parity": parity, "roofline": roofline, "cpu_baseline": cpu,
           "fitness_checksum": float(np.nansum(fit_host)), "nonzero_fitness_frac": float((fit_host > 0).mean())}
    if precision is None:
        rec["config"]["precision_profile"] = int(os.environ.get("EIG_PRECISION", "0") or 0)
    if precision is not None:
        rec["config"]["precision_profile"] = precision
        rec["dtype"] = ("fp32-grade (3 split-fp16 products per MAC)", "mixed: single fp16 product in PredNet layers 2+3, 3 products in layer 1",
                        "single fp16 product per MAC, fp32 accumulate")[precision]
        rec["roofline"]["mma_passes"] = (3, "3 (layer 1) / 1 (layers 2, 3)", 1)[precision]
        rec["roofline"]["frac_of_3pass_ceiling"] = None if precision else rec["roofline"]["frac_of_3pass_ceiling"]
    if gathered_check is not None:
        rec["gathered_check"] = gathered_check
    if e2e_genomes is not None:
        rec["e2e_from_genomes"] = e2e_genomes
    return rec


def time_from_genomes(eng, genomes, cfg, c_dim, structure, steps):
    """Wall clock from genome OBJECTS to the host fitness vector (flatten + pack + H2D + kernels + D2H), the level at which
    `get_fitnesses_neat` is called: cold = every genome flattened, warm = every genome a program-cache hit."""
    import torch
    from evolutionary_illusion_generator_b200 import genome as G, runtime
    n_out = c_dim if c_dim > 1 else 1
    items = list(enumerate(genomes))
    out = {}
    for mode in ("cold", "warm"):
        cache = G.ProgramCache()

        def flatten(gid, g):
            return cache.get(gid, g, cfg, n_out)

        def once():
            if mode == "cold":
                cache._entries.clear()
            return runtime.evaluate_genomes(eng, items, flatten, structure)

        once(); once()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            once()
        dt = time.perf_counter() - t0
        out[mode] = {"value": len(items) * steps / dt, "unit": "evals/s", "ms_per_step": 1e3 * dt / steps}
    out["path"] = "runtime.evaluate_genomes: genome objects -> flatten (program cache) -> pack -> pinned H2D -> eig_eval -> D2H"
    return out


def run_ours(args):
    import torch
    wl = args.workload
    world = int(os.environ.get("WORLD_SIZE", "1"))
    scaling = args.scaling
    wanted = [a for a in args.also.split(",") if a] if args.also != "auto" else None
    if wanted is None:
        wanted = []
        if wl == "c3" and scaling == "weak" and not args.pop:
            wanted.append("c2")
            if world > 1:
                wanted.append("c3_strong")
            if world == 8:
                wanted += ["c4", "c5"]
            if world == 1 and args.conv == "tc":
                wanted += ["c3_precision1", "c3_precision2"]
    # The oracle side of every parity gate runs on rank 0 BEFORE the process group exists: NCCL's initialisation narrows
    # the CPU affinity its caller's later OpenMP workers inherit (measured: the same 4 genomes took 1.6 s at N = 1 and
    # 218 s at N = 2 when the oracle ran after init_process_group).
    refs = {}
    if int(os.environ.get("RANK", "0")) == 0 and args.parity > 0:
        refs[wl] = oracle_fitness(wl, 0, args.parity, os.cpu_count() or 1)
        for name in wanted:
            if name in ("c2", "c4"):
                refs[name] = oracle_fitness(name, 0, 2, os.cpu_count() or 1)
    D = Dist()
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device=D.dev)   # > 126 MB L2
    pop = args.pop or WORKLOADS[wl][6]
    if scaling == "strong":   # SURVEY.md §8 d sweep (i): the workload's population split over the ranks
        pop = -(-pop // D.world)

    def ref_for(name):
        if args.parity <= 0 or (name not in refs and D.rank == 0):
            return None
        return refs.get(name, (None, 0.0))      # ranks > 0 only need to know that a gate runs

    main = measure(D, wl, pop, args.steps, args.warmup, scaling, args.conv, flush, True, ref_for(wl),
                   0 if args.no_cpu_baseline else (args.ref_sample or default_ref_sample(wl) * 4), True)
    also = []
    for name in wanted:
        if name == "c3_strong":
            rec = measure(D, "c3", -(-WORKLOADS["c3"][6] // D.world), args.steps, args.warmup, "strong", args.conv, flush,
                          False, None, 0, False)
        elif name.startswith("c3_precision"):   # the opt-in precision profiles, same population and gate as the contract line
            rec = measure(D, "c3", WORKLOADS["c3"][6], args.steps, args.warmup, "weak", args.conv, flush, False,
                          ref_for("c3") if wl == "c3" else None, 0, False, precision=int(name[-1]))
        else:
            short = name == "c5"      # 1.4 s per step: keep the sub-record to a few seconds
            rec = measure(D, name, WORKLOADS[name][6], 3 if short else args.steps, 3 if short else args.warmup, "weak",
                          args.conv, flush, False, ref_for(name) if name in ("c2", "c4") else None, 0, False)
        if rec is not None:
            rec["name"] = name
            also.append(rec)
    if D.rank == 0:
        main["also"] = also
        print(json.dumps(main))
    D.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--conv", default=os.environ.get("EIG_BENCH_CONV", "auto"), choices=["auto", "simt", "tc"])
    ap.add_argument("--pop", type=int, default=0, help="genomes per GPU (default: the workload's)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: fixed genomes per GPU (default, the driver's contract); strong: the population is split over the GPUs")
    ap.add_argument("--also", default="auto", help="comma list of sub-records (c2,c3_strong,c4,c5), '' for none; auto = by N")
    ap.add_argument("--parity", type=int, default=4, help="genomes of the pre-timing parity gate (0 = off)")
    ap.add_argument("--ref-sample", type=int, default=0, help="genomes per CPU-baseline sample (default: by workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.conv == "auto":
        args.conv = "tc" if tc_compiled() else "simt"
    run_ours(args)


def tc_compiled():
    """True when libeig.so carries the tcgen05 convolution (probed through the C ABI on a live context)."""
    try:
        import torch
        from evolutionary_illusion_generator_b200 import _lib, engine as E
        if not torch.cuda.is_available():
            return False
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        e = E.Engine(64, 64, (1, 4, 8, 8), 2)
        try:
            e.set_conv_mode(_lib.CONV_TC)
            ok = True
        except _lib.EigError:
            ok = False
        e.close()
        return ok
    except Exception:
        return False


if __name__ == "__main__":
    main()
