#!/usr/bin/env python
"""bench.py - genome fitness evaluations per second on N B200s (BASELINE.json metric).  `--workload c4|c5`: the per-GPU shares of the 320x240 / 512x512 configs (not 160x120; informational).

One "step" = one pass of the hot path (CPPN render -> PredNet 20+1 steps -> Shi-Tomasi/LK flow -> score) over
one synthetic population shard per GPU.  Default workload = BASELINE.json configs[1] ("c2": pop 32,
neat_configs/circles_bw.txt, 160x120 gray, PredNet channels 1,16,32,64); `--workload c3` is configs[2]
(circles.txt colour, 3,48,96,192).  Weak scaling: every GPU owns `pop` genomes, the fitness slices are
all-gathered (NCCL) inside every step.

  value : evals/s with the flattened genomes already resident in HBM; device time from CUDA events per step
          (L2 flushed between steps, untimed), max over ranks.
  e2e   : the same through the host entry point `eig_eval_host` (what get_fitnesses_neat calls): pinned host
          genome blob -> H2D -> kernels -> D2H fitness, wall clock with a device sync, max over ranks.
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
STRUCTURE_NAMES = {0: "Bands", 1: "Circles", 2: "Free", 3: "CirclesFree"}


def tc_kernel_macs(ch):
    """Algorithmic MACs per layer-0 pixel per PredNet step that the tcgen05 kernel covers (SURVEY.md §8 a-7 counts):
    everything except ConvA1, ConvP0 and the E0/h0 taps of ConvLSTM0, which run on the layer-0 SIMT kernels.
    ConvLSTM0's up-sampled-R1 taps are counted at their reference cost 9*C1*4*C0 (the kernel evaluates them folded to
    half resolution)."""
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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_population(preset, c_dim, n, start):
    from evolutionary_illusion_generator_b200 import genome as G
    n_out = G.NEAT_PRESETS[preset]["num_outputs"]
    cfg = G.make_config(2, n_out)
    pop = [G.synthetic_genome(preset, start + i) for i in range(n)]
    progs = [G.flatten_genome(g, cfg, n_outputs=c_dim if c_dim > 1 else 1) for g in pop]
    return cfg, pop, progs


def oracle_evals_per_s(workload, n_sample, threads, flow_impl="cv2"):
    """The reference's CPU path (oracle port) on `n_sample` genomes of the workload."""
    import torch
    from evolutionary_illusion_generator_b200 import weights as W
    from oracle import pipeline as OPL
    preset, c_dim, ch, w, h, structure, _, _ = WORKLOADS[workload]
    torch.set_num_threads(threads)
    wts = W.synthetic_predictor_weights(w, h, ch, seed=0)
    cfg, pop, _ = build_population(preset, c_dim, n_sample, 0)
    gc = cfg.genome_config
    t0 = time.perf_counter()
    OPL.evaluate_population(pop, gc.input_keys, gc.output_keys, structure, wts, w, h, ch, c_dim, flow_impl=flow_impl)
    dt = time.perf_counter() - t0
    return n_sample / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    preset, c_dim, ch, w, h, structure, pop, _ = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    n_sample = args.ref_sample
    for _ in range(max(args.warmup, 1) if args.warmup < 2 else 1):
        oracle_evals_per_s(args.workload, 1, cores)
    vals, secs = [], 0.0
    for _ in range(args.steps):
        v, dt = oracle_evals_per_s(args.workload, n_sample, cores)
        vals.append(v); secs += dt
    value = args.steps * n_sample / secs
    sample = ("%d genomes per step of workload %s (render + 22 PredNet forwards + cv2 LK + scoring, in memory), "
              "torch-CPU fp32 oracle port of the Chainer path, %d threads" % (n_sample, args.workload, cores))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": args.workload, "pop_per_step": n_sample, "resolution": "%dx%d" % (w, h),
                       "channels": list(ch), "neat_config": preset},
            "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from evolutionary_illusion_generator_b200 import _lib, engine as E, genome as G, weights as W

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    preset, c_dim, ch, w, h, structure, pop, macs = WORKLOADS[args.workload]
    if args.pop:
        pop = args.pop
    if args.scaling == "strong":   # SURVEY.md §8 d sweep (i): the workload's population split over the ranks
        pop = -(-pop // world)
    dev = torch.device("cuda", local)
    eng = E.Engine(w, h, ch, pop, device=local)
    eng.set_conv_mode(_lib.CONV_TC if args.conv == "tc" else _lib.CONV_SIMT)
    eng.set_grid(structure)
    eng.load_weights(W.synthetic_predictor_weights(w, h, ch, seed=0))
    _, _, progs = build_population(preset, c_dim, pop, rank * pop)
    blob, offsets, max_slots = G.pack_population(progs)
    resident = eng.upload_programs(progs)
    fit_dev = torch.empty((pop,), dtype=torch.float64, device=dev)
    gathered = torch.empty((world * pop,), dtype=torch.float64, device=dev)
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)   # > 126 MB L2
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_resident()
        flush.zero_()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.lib.eig_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in ev:
        a.record()
        step_resident()
        b.record()
        flush.zero_()          # L2 flush between timed steps (not inside any event pair)
    barrier()
    launches = eng.lib.eig_launch_count() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    # end-to-end through the host entry point
    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    tmax = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(tmax[0]), float(tmax[1])
    fit_host = fit_dev.cpu().numpy()

    # instrumented pass: per-kernel-class device time of the same step (rank 0 reports)
    import ctypes as C
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

    if rank == 0:
        peaks = measured_peaks()
        total = world * pop * args.steps
        value = total / (dev_ms / 1e3)
        e2e = total / (e2e_ms / 1e3)
        flop_step = 2.0 * (tc_kernel_macs(ch) * USEFUL_STEPS - tc_dead_macs(ch)) * w * h * pop   # algorithmic FLOP of the tcgen05 launches, one GPU
        conv_ms = cls_ms["conv_simt"] + cls_ms["conv_tcgen05"]
        conv_n = cls_n["conv_simt"] + cls_n["conv_tcgen05"]
        achieved = flop_step / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r1", "tc_traffic.json")
        if args.conv == "tc" and os.path.isfile(tpath) and not args.pop:
            traffic = json.load(open(tpath)).get(args.workload, {}).get("mean_dram_bytes_per_launch")
        roofline = {"bound": "tensor", "kernel": "conv3x3_tc_kernel (%s)" % ("tcgen05 cta_group::2, 3-pass split fp16" if args.conv == "tc" else "fp32 SIMT"),
                    "achieved": achieved, "peak": peaks["tf"], "unit": "TFLOP/s", "frac": achieved / peaks["tf"],
                    "traffic": traffic, "traffic_source": "ncu --set full, profiles/r1/tc_traffic.json (mean over the 8 launch shapes of one PredNet step)" if traffic else None,
                    "peak_source": peaks["source"] + " bf16 dense (sustained)",
                    "flop_per_launch": flop_step / max(conv_n, 1), "launches_per_step": conv_n,
                    "avg_launch_us": 1e3 * conv_ms / max(conv_n, 1),
                    "mma_passes": 3 if args.conv == "tc" else None,
                    "frac_of_3pass_ceiling": 3.0 * achieved / peaks["tf"] if args.conv == "tc" else None,
                    "class_ms_per_step": cls_ms, "class_launches_per_step": cls_n,
                    "whole_path_gflop_per_genome": 2.0 * macs * w * h * USEFUL_STEPS / 1e9,
                    "timing": "per-launch CUDA events of an instrumented pass of the same step (library launch instrumentation; "
                              "graph replay, side-stream overlap and programmatic dependent launch are off in that pass, "
                              "so the class times are upper bounds of their share of ms_per_step)",
                    "note": "achieved counts each algorithmic MAC once; fp32-grade accuracy needs 3 fp16 MMAs per MAC "
                            "(hi*hi + hi*lo + lo*hi), so 1/3 of the dense 16-bit peak is the ceiling of this kernel"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_s = args.ref_sample
            oracle_evals_per_s(args.workload, 1, cores)
            v, dt = oracle_evals_per_s(args.workload, n_s, cores)
            cpu = {"value": v, "unit": "evals/s", "cores": cores, "kind": "port",
                   "sample": "%d genomes of workload %s through the oracle (torch-CPU fp32 PredNet port, cv2 LK), "
                             "%.1f s" % (n_s, args.workload, dt)}
        line = {"metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "fp32 (3-pass split-fp16 tcgen05, fp32 accumulate)" if args.conv == "tc" else "fp32",
                "data": "synthetic",
                "config": {"workload": args.workload, "pop_per_gpu": pop, "global_pop": world * pop,
                           "resolution": "%dx%d" % (w, h), "channels": list(ch), "neat_config": preset,
                           "structure": STRUCTURE_NAMES[structure], "prednet_steps": USEFUL_STEPS, "conv": args.conv,
                           "weights": "synthetic_predictor_weights seed 0 (LeCun-normal, layer 0 shaped as an error integrator)", "l2": "flushed between steps (256 MiB memset, untimed)",
                           "parallelism": "genome-sharded dp%d + 1 all-gather/step" % world},
                "e2e": {"value": e2e, "unit": "evals/s", "h2d_bytes_per_step": int(blob.nbytes + offsets.nbytes),
                        "d2h_bytes_per_step": int(8 * pop), "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "fitness_checksum": float(np.nansum(fit_host)), "nonzero_fitness_frac": float((fit_host > 0).mean())}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--conv", default=os.environ.get("EIG_BENCH_CONV", "auto"), choices=["auto", "simt", "tc"])
    ap.add_argument("--pop", type=int, default=0, help="genomes per GPU (default: the workload's)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: fixed genomes per GPU (default, the driver's contract); strong: the population is split over the GPUs")
    ap.add_argument("--ref-sample", type=int, default=8, help="genomes per CPU-baseline sample")
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
