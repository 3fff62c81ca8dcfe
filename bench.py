#!/usr/bin/env python
"""bench.py — benchmark of the per-bin integration hot path (BASELINE.json metric: integrand evals/s, 1024^2 bins MC+CV, % of roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c2b|c2k16|c3|c4|c5] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input = ONE call of the reference-facing integrator:

  c2   monte_carlo_per_bin_parallel(64, seed) over 1024x1024 bins of shade4<64> (BASELINE configs[1]) — the "MC" half of the headline
  c4   integrator_crespo2021(65536, 64, seed) over 1024x1024 bins of shade5<64> (BASELINE configs[3]) — the "CV" half: region table
       (batched generation) + control variate + residual Monte Carlo, bins sharded over the N GPUs
  c3   integrator_adaptive_iterations(nested(boole,simpson), size/relative 1e-5, 10^6) over 512x512 bins of smooth_edge2 (configs[2])
  c5   monte_carlo_per_bin_parallel(256, seed) over 2048x2048 bins of the Russian-roulette walk, RangeInfinite (configs[4])

Prints ONE JSON line (rank 0).  With no --workload the line is C2's (the configuration the metric is quoted on) and carries C4's full
record — same keys — under "cv" (the second headline), so one default run measures both halves of "MC+CV".

Keys (per record): `value` = whole-job throughput with the bins resident in HBM, CUDA events per step on the library's stream, L2
flushed between steps, max over ranks; `sustained` = the same step repeated back to back for >= 2 s (median SM clock over it); `value`
is the burst figure only when the two agree within 3 %, else the sustained one (`value_source` says which).  `e2e` = the same call with
HOST bins through the C ABI (device->host copy of the bins and the host-side '+=' / '=' inside the timed region).  `roofline`: the
non-tensor FP32 roofline for c2/c4/c5 (bound "fp32": this path has no tensor-core work and C2 moves 4 B per 64 evaluations, so neither
of the contract's "hbm"/"tensor" bounds describes it), HBM for c3; `achieved` = algorithmic flops (bytes) per step, SURVEY.md §8(d)'s
per-unit figures x the units of a step, / the measured step time.  `cpu_baseline` / `--impl reference`: the UNMODIFIED reference on this
box's host cores (oracle/_ref, multi-threaded through a std::thread parallel-STL back end where upstream uses TBB), steps sized >= 10 s
so that the 250 ms quantum of its progress logger (reference src/foreach.h:48-55) stays below 3 %.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# F_alg per unit, SURVEY.md §8(d)
F_PAIR, F_APPROX, F_F5 = 876, 846, 158
WORKLOADS = {
    "c2": dict(desc="per-bin stratified MC, 1024x1024 bins, 64 spp, shade4<64> 4D, fp32 (BASELINE configs[1])", integrand="shade4_64", res=[1024, 1024], spp=64,
               path="mc_per_bin_parallel", flops=155, kind="mc"),
    "c2b": dict(desc="integrator_per_bin_parallel(monte_carlo(64)), 1024x1024 bins, shade4<64>", integrand="shade4_64", res=[1024, 1024], spp=64,
                path="per_bin_parallel_mc", flops=155, kind="mc"),
    "c2k16": dict(desc="per-bin stratified MC, 1024x1024 bins, 64 spp, shade4<16>", integrand="shade4_16", res=[1024, 1024], spp=64, path="mc_per_bin_parallel", flops=59, kind="mc"),
    "c3": dict(desc="nested Newton-Cotes (boole/simpson) adaptive refinement, 10^6 iterations, 512x512 bins, smooth_edge2 (BASELINE configs[2])", integrand="smooth_edge2",
               res=[512, 512], iterations=1000000, kind="nc"),
    "c4": dict(desc="integrator_crespo2021(65536, 64): control variates + residual MC, 1024x1024 bins, shade5<64> 5D (BASELINE configs[3])", integrand="shade5_64",
               res=[1024, 1024], spp=64, iterations=65536, kind="cv"),
    "c5": dict(desc="range_infinite random walk with Russian roulette, 2048x2048 bins, 256 spp (BASELINE configs[4])", integrand="walk", res=[2048, 2048], spp=256,
               path="mc_per_bin_parallel_inf", kind="walk"),
}
METRIC = {"c2": "integrand evals/sec (per-bin MC, 1024x1024 bins x 64 spp, shade4<64>)",
          "c4": "integrand evals/sec (control variates + residual MC, integrator_crespo2021(65536,64), 1024x1024 bins, shade5<64>)",
          "c3": "regions/sec (nested Newton-Cotes adaptive refinement, 10^6 iterations, 512x512 bins)",
          "c5": "paths/sec (range_infinite random walk, 2048x2048 bins x 256 spp)"}
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the ncu --set full captures under profiles/
NCU_TRAFFIC_BYTES = {"c2": 36864 + 4194304, "c5": 16922112, "c4": 100966400 + 949487616}      # dram read + write of the dominant kernel per launch (profiles/ncu_*)
RNG_NOTE = {"mc": "xoshiro128++ stream per (bin, lane sub-stream), state = Philox4x32-10(key=seed, counter=(bin, sub-stream)); 20 words per group of 8 samples: 24-bit fields for "
                  "the free dimensions, 16-bit fields inside a bin of a >=256-bin axis (VB200_MC_RNG_PHILOX selects Philox4x32-10 for every draw: see config.philox_value)",
            "walk": "Philox4x32-10, counter (bin, sample, block): one block of four 24-bit elements per lane and loop iteration, first-round products cached (18 multiplies per block)",
            "cv": "Philox4x32-10 keyed by (seed; bin, sample): region choice + in-region point"}


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region through NVML (same data as the nvidia-smi
    clocks line of B200_PROFILING.md, but at ~1 kHz so that a millisecond-scale timed region still gets samples)."""

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.stop_flag, self.err = index, [], 0, False, None
        self.sm_max = None
        self.marks = {}

    def _loop(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if visible:
                try:
                    idx = int(visible.split(",")[self.index])
                except Exception:
                    pass
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            while not self.stop_flag:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.reasons |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                time.sleep(0.001)
        except Exception as ex:
            self.err = str(ex)

    def start(self):
        self.t = threading.Thread(target=self._loop, daemon=True); self.t.start()

    def mark(self, name):
        self.marks[name] = len(self.sm)

    def median_between(self, a, b):
        s = self.sm[self.marks.get(a, 0):self.marks.get(b, len(self.sm))]
        return float(np.median(s)) if s else None

    def stop(self):
        self.stop_flag = True
        self.t.join(timeout=2)
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}
        reasons = sorted(n for bit, n in names.items() if self.reasons & bit)
        out = {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max, "reasons": reasons, "samples": len(self.sm)}
        if self.err:
            out["error"] = self.err
        return out


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---- the reference's CPU path ------------------------------------------------------------------------------------------------
QUANTUM_NOTE = "wall clock of the reference call, which ends on a 250 ms tick of its progress-logger thread (reference src/foreach.h:48-55): <= 2.5 % of a >= 10 s step"


def _ref_libs():
    """(multi-threaded reference, serial reference or port, kind).  oracle/_ref/libviltrum_ref_mt.so = the unmodified reference with a
    std::thread back end under its std::for_each(par_unseq) loops; the port (oracle/liboracle.so) only where _ref was never built."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    mt = pyoracle.load("reference-mt") if (pyoracle.available("reference-mt") or os.path.isdir(pyoracle.REFERENCE_ROOT)) else None
    if pyoracle.available("reference") or os.path.isdir(pyoracle.REFERENCE_ROOT):
        return mt, pyoracle.load("reference"), "reference"
    return None, pyoracle.load("port"), "port"


def cpu_step(workload, seed, budget_s=12.0):
    """One timed call of the reference's CPU implementation of `workload` -> (units, seconds, cores, sample text, extra)."""
    w = WORKLOADS[workload]
    mt, serial, kind = _ref_libs()
    T = cpu_threads()
    if w["kind"] in ("mc", "walk"):
        # the full bin grid, stacked M times along the last bin dimension so that one call lasts >= ~10 s: every bin still takes spp samples of
        # its own freshly seeded generator (the per-bin cost structure of the workload), the logger quantum drops below 3 %
        rate_guess = {"mc": 6.5e6, "walk": 11e6}[w["kind"]] * T      # evals/s per thread measured on the pool's Xeon hosts
        units1 = w["res"][0] * w["res"][1] * w["spp"]
        M = max(1, int(np.ceil(budget_s * rate_guess / units1)))
        res = [w["res"][0], w["res"][1] * M]
        inf = w["kind"] == "walk"
        t0 = time.perf_counter()
        if mt is not None:
            mt.set_threads(T)
            if inf:
                bins = mt.mc_per_bin_parallel_inf(w["integrand"], res, w["spp"], seed)
            else:
                bins = getattr(mt, w["path"])(w["integrand"], res, [0.0] * 4, [1.0] * 4, w["spp"], seed)
            how = f"unmodified reference, its par_unseq loops on {T} std::threads"
        else:
            rmin, rmax = ((), ()) if inf else ([0.0] * 4, [1.0] * 4)
            bins = serial.mt_per_bin(w["path"], w["integrand"], res, w["spp"], seed, T, rmin, rmax)
            how = f"bin grid slabbed over {T} threads, each an independent single-threaded call"
        dt = time.perf_counter() - t0
        assert np.all(np.isfinite(bins))
        units = units1 * M
        return units, dt, T, (f"{M} stacked frames of the {w['res'][0]}x{w['res'][1]}-bin x {w['spp']} spp workload per call ({units/1e6:.0f} M evals, grid {res[0]}x{res[1]}); {how}; "
                              + QUANTUM_NOTE), {}
    if w["kind"] == "nc":
        # the whole job: generation is a serial heap loop upstream (regions-generator-adaptive-heap.h:18-45), accumulation sequential
        t0 = time.perf_counter()
        bins, _ = serial.adaptive_iterations(w["integrand"], "boole_simpson", "size_relative", w["iterations"], w["res"], [0.0, 0.0], [1.0, 1.0], 1e-5)
        dt = time.perf_counter() - t0
        ph = serial.phase_times() if hasattr(serial, "phase_times") else None
        return w["iterations"], dt, 1, "the full job on 1 thread (generation and accumulation are serial upstream)", {"generation_s": ph[0] if ph else None}
    if w["kind"] == "cv":
        # The full job needs ~1e9 (bin, region) pointers = 8+ GB of per-bin vectors in the reference (mutexed-tensor-vector.h) and ~15 CPU-minutes:
        # sampled on a coarser bin grid over the same region table (65 536 iterations) and scaled by the bin count — the per-bin cost is set
        # by the regions that touch a bin (967.6 at 1024^2; 5-D regions are ~0.11 wide, so it barely depends on the bin size) and by spp.
        side = 256
        lib = mt if mt is not None else serial
        if mt is not None:
            mt.set_threads(T)
        t0 = time.perf_counter()
        bins, _ = lib.crespo2021(w["integrand"], w["iterations"], w["spp"], seed, [side, side], [0.0] * 5, [1.0] * 5)
        dt = time.perf_counter() - t0
        ph = lib.phase_times() if hasattr(lib, "phase_times") else (0.0, dt)
        t_gen, t_cv = ph[0], ph[1] - ph[0]
        scale = (w["res"][0] * w["res"][1]) / float(side * side)
        full_s = t_gen + t_cv * scale
        units = w["res"][0] * w["res"][1] * w["spp"]
        cores = T if mt is not None else 1
        return units, full_s, cores, (f"integrator_crespo2021(65536, 64) over {side}x{side} bins measured: generation {t_gen:.2f} s (serial upstream) + stratification/CV/residual "
                                     f"{t_cv:.2f} s on {cores} threads; scaled to 1024x1024 bins: {t_gen:.2f} + {t_cv:.2f} x {scale:.0f} = {full_s:.1f} s per step (the full job would need "
                                     f">= 8 GB of per-bin region lists in the reference); " + QUANTUM_NOTE), {"measured_s": dt, "generation_s": t_gen}
    raise ValueError(workload)


def run_cpu(workload, steps, warmup, wall_budget_s):
    """steps timed reference calls (fewer if the wall budget runs out; at least one) after `warmup` untimed ones (at most one is worth its time)"""
    _, _, kind = _ref_libs()
    t_start = time.perf_counter()
    for i in range(min(warmup, 1)):
        cpu_step(workload, i, budget_s=2.0 if WORKLOADS[workload]["kind"] in ("mc", "walk") else 0)
    secs, done, last = [], 0, None
    for i in range(steps):
        last = cpu_step(workload, 100 + i)
        secs.append(last[1]); done += 1
        if time.perf_counter() - t_start + last[1] > wall_budget_s:
            break
    units, _, cores, sample, extra = last
    sec = float(np.mean(secs))
    unit = "regions/s" if WORKLOADS[workload]["kind"] == "nc" else ("paths/s" if WORKLOADS[workload]["kind"] == "walk" else "evals/s")
    cb = dict(value=units / sec, unit=unit, cores=cores, kind=kind, sample=sample + f"; mean of {done} step(s)", seconds_per_step=sec, **extra)
    return cb, sec, units, done


def config_of(workload, world=1, scaling="weak"):
    w = WORKLOADS[workload]
    cfg = {"workload": w["desc"], "integrand": w["integrand"], "bins": w["res"]}
    if "spp" in w:
        cfg["spp"] = w["spp"]
    if "iterations" in w:
        cfg["iterations"] = w["iterations"]
    return cfg


# ---- GPU arm -------------------------------------------------------------------------------------------------------------------
def mean_walk_bounces():
    """E[bounces] of the walk integrand over the image: alb/(1-alb), alb = .4 + .5*(4 px (1-px))*(.25 + .75 py)"""
    x = (np.arange(2048) + 0.5) / 2048.0
    alb = 0.4 + 0.5 * (4 * x * (1 - x))[None, :] * (0.25 + 0.75 * x)[:, None]
    return float(np.mean(alb / (1 - alb)))


class Bench:
    def __init__(self, workload, args, rank, local_rank, world, ctx, stream, sampler, flush):
        import torch
        self.torch, self.w, self.name, self.args = torch, WORKLOADS[workload], workload, args
        self.rank, self.local_rank, self.world, self.ctx, self.stream, self.sampler, self.flush = rank, local_rank, world, ctx, stream, sampler, flush
        w = self.w
        from viltrum_b200 import Range, RangeInfinite, shard_for_rank, _capi
        self.capi = _capi
        # scaling: c2 weak by default (every rank owns a full-size slab of a grid N times as tall: independent bins, no collective), c5 weak as replicas, c3/c4 strong
        # (BASELINE configs[3]: "bins sharded over 8 GPUs"); --scaling overrides
        self.scaling = args.scaling or ("weak" if w["kind"] in ("mc", "walk") else "strong")
        res = list(w["res"])
        # region-based workloads under --scaling weak: N replicas, every rank integrates the whole grid with its own seed (the region table does not
        # shard, so "more bins" would change the regions per bin); the per-bin samplers stack the ranks' slabs into one taller grid instead
        # (the walk integrand's path length depends on where a bin sits in the image: slabs of a taller grid would not be equal work, so it is replicated too)
        self.replicas = self.scaling == "weak" and w["kind"] in ("nc", "cv", "walk") and world > 1
        self.gres = [res[0], res[1] * world] if (self.scaling == "weak" and not self.replicas) else res
        self.shard = (0, self.gres[0] * self.gres[1]) if self.replicas else shard_for_rank(self.gres, rank, world)
        self.nb_local = self.shard[1] - self.shard[0]
        self.nb_global = self.gres[0] * self.gres[1]
        d = {"mc": 4, "cv": 5, "nc": 2}.get(w["kind"])
        self.rng = RangeInfinite() if w["kind"] == "walk" else Range([0.0] * d, [1.0] * d)
        self.d_bins = torch.zeros(self.nb_global, dtype=torch.float32, device="cuda")
        self.h_bins = np.zeros(self.nb_global, np.float32)
        self.d_nreg = torch.zeros(self.nb_global, dtype=torch.int32, device="cuda") if w["kind"] == "cv" else None
        self.generator = "xoshiro"
        self.ktimer = False

    # one step = one integrator call
    def step(self, bins, seed):
        w, ctx = self.w, self.ctx
        if w["kind"] == "mc":
            ctx.mc_per_bin(w["integrand"], bins, self.gres, self.rng, w["spp"], seed, self.capi.MC_PER_BIN if w["path"] == "mc_per_bin_parallel" else self.capi.PER_BIN_MC,
                           shard=self.shard, generator=self.generator)
        elif w["kind"] == "walk":
            ctx.mc_per_bin_inf(w["integrand"], bins, self.gres, self.rng, w["spp"], seed + 7919 * (self.rank if self.replicas else 0), shard=self.shard)
        elif w["kind"] == "nc":
            regs = ctx.regions_generate_adaptive(w["integrand"], self.rng, "boole_simpson", "size", "relative", w["iterations"], 1e-5, batch=self.batch, exact=True)
            regs.integrate_bins(bins, self.gres, self.rng, shard=self.shard)
            regs.free()
        elif w["kind"] == "cv":
            # default: every rank generates the (identical, deterministic) table — generation cannot be sharded, so T_gen + T_cv/N beats
            # T_gen + T_broadcast + T_cv/N.  --table broadcast: rank 0 generates, vb200_regions_broadcast (NCCL over NVLink) hands it out.
            if self.args.table == "broadcast" and self.world > 1:
                regs = ctx.regions_generate_adaptive(w["integrand"], self.rng, "simpson_trapezoidal", "size", "relative", w["iterations"], 1e-5, batch=self.batch,
                                                     exact=True) if self.rank == 0 else None
                regs = ctx.regions_broadcast(regs, 0)
            else:
                regs = ctx.regions_generate_adaptive(w["integrand"], self.rng, "simpson_trapezoidal", "size", "relative", w["iterations"], 1e-5, batch=self.batch, exact=True)
            regs.cv_integrate(w["integrand"], bins, self.gres, self.rng, w["spp"], seed + 7919 * (self.rank if self.replicas else 0), shard=self.shard,
                              nregions=self.d_nreg if (self.d_nreg is not None and not isinstance(bins, np.ndarray)) else None)
            regs.free()

    batch = 0       # region generators: 0 = batched top-k refinement (throughput mode); 1 = the reference's exact greedy order

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.world > 1:
            import torch.distributed as dist
            t = self.torch.tensor([ms], dtype=self.torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        return ms

    def timed(self, bins, steps, warmup):
        """per-step CUDA events on the library's stream, L2 flushed (untimed) between steps; mean over the steps, max over ranks"""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            for i in range(warmup):
                self.step(bins, i)
            self.barrier()
            if self.ktimer:
                self.ctx.kernel_timer_read()                                       # drop the warm-up launches
            evs = []
            for i in range(steps):
                self.flush.zero_()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(self.stream); self.step(bins, 1000 + i); e1.record(self.stream)
                evs.append((e0, e1))
            self.barrier()
        return self.max_over_ranks(sum(a.elapsed_time(b) for a, b in evs) / steps)

    def sustained(self, bins, ms_burst, seconds):
        """the same step back to back (no flush, one event pair around the whole run) for >= `seconds`"""
        torch = self.torch
        n = max(3, int(np.ceil(seconds * 1e3 / max(ms_burst, 1e-3))))
        with torch.cuda.stream(self.stream):
            self.barrier()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            if self.sampler:
                self.sampler.mark(self.name + "_s0")
            e0.record(self.stream)
            for i in range(n):
                self.step(bins, 5000 + i)
            e1.record(self.stream)
            self.barrier()
            if self.sampler:
                self.sampler.mark(self.name + "_s1")
        ms = self.max_over_ranks(e0.elapsed_time(e1) / n)
        return ms, n

    def units(self):
        w = self.w
        rep = self.world if self.replicas else 1
        if w["kind"] == "nc":
            return w["iterations"] * rep
        return self.nb_global * w["spp"] * rep

    def run(self, peaks, fp32_peak, fma_peak, sm_max):
        args, w, ctx = self.args, self.w, self.ctx
        unit = "regions/s" if w["kind"] == "nc" else ("paths/s" if w["kind"] == "walk" else "evals/s")
        l0 = ctx.launch_count
        if w["kind"] == "cv":                                                      # event pairs around the heaviest kernel's launches (vb200_kernel_timer)
            ctx.kernel_timer(True); self.ktimer = True
        ms_dev = self.timed(self.d_bins, args.steps, args.warmup)
        k_ms, k_launches = (0.0, 0)
        if self.ktimer:
            k_ms, k_launches = ctx.kernel_timer_read(); ctx.kernel_timer(False); self.ktimer = False
        launches = ctx.launch_count - l0 - 0
        launches_per_step = launches / float(args.steps + args.warmup)
        ms_sus, n_sus = self.sustained(self.d_bins, ms_dev, args.sustain)
        ms_e2e = self.timed(self.h_bins, args.steps, max(3, args.warmup))          # caller's bins in ordinary pageable memory
        rec_extra = {}
        if w["kind"] in ("mc", "walk"):
            ctx.host_register(self.h_bins)                                         # optional: caller's bins pinned + mapped once, outside the timed loop
            ms_pin = self.timed(self.h_bins, args.steps, max(3, args.warmup))
            ctx.host_unregister(self.h_bins)
            rec_extra["pinned_value"] = self.units() / (ms_pin * 1e-3); rec_extra["pinned_ms_per_step"] = ms_pin
        philox = None
        if w["kind"] == "mc":                                                      # the selectable pure-Philox stream, same launch
            self.generator = "philox"
            philox = self.units() / (self.timed(self.d_bins, args.steps, args.warmup) * 1e-3)
            self.generator = "xoshiro"
        exact = None
        if w["kind"] in ("nc", "cv") and self.world == 1 and (args.exact or w["kind"] == "cv"):
            self.batch = 1                                                         # the reference's greedy split order, bit-identical region list
            self.step(self.d_bins, 1); ctx.synchronize()
            t0 = time.perf_counter(); self.step(self.d_bins, 2); ctx.synchronize()
            exact = {"ms_per_step": (time.perf_counter() - t0) * 1e3, "note": "batch = 1: identical region list to the reference (parity mode); one call, wall clock"}
            self.batch = 0
        units = self.units()
        burst, sus = units / (ms_dev * 1e-3), units / (ms_sus * 1e-3)
        agree = abs(burst - sus) <= 0.03 * sus
        value, ms_value = (burst, ms_dev) if agree else (sus, ms_sus)
        sus_clock = self.sampler.median_between(self.name + "_s0", self.name + "_s1") if self.sampler else None

        # ---- roofline (per GPU: this rank's share of the step) ----
        per_gpu_s = ms_value * 1e-3
        roof = None
        if w["kind"] == "mc":
            achieved = self.nb_local * w["spp"] * w["flops"] / per_gpu_s / 1e12
            roof = {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak, "traffic": NCU_TRAFFIC_BYTES.get(self.name),
                    "flops_per_eval": w["flops"], "evals_per_launch": self.nb_local * w["spp"], "hbm_gbs_achieved": self.nb_local * 4 / per_gpu_s / 1e9}
        elif w["kind"] == "walk":
            b = mean_walk_bounces()
            fl = 9 + 6 * b
            achieved = self.nb_local * w["spp"] * fl / per_gpu_s / 1e12
            roof = {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak, "traffic": NCU_TRAFFIC_BYTES.get(self.name),
                    "flops_per_path": fl, "mean_bounces": b, "randoms_per_path": 3 + 2 * b, "randoms_per_s": self.nb_local * w["spp"] * (3 + 2 * b) / per_gpu_s,
                    "issue_slot_pct": 68.0, "issue_slot_source": "profiles/ncu_c5_walk_window_r1d.txt (smsp__issue_active)",
                    "note": "issue bound by the generator and the per-lane roulette (~6 flops per random number): the FP32 fraction is low by construction (SURVEY.md §8d)"}
        elif w["kind"] == "cv":
            pairs = float(self.d_nreg.double().sum().item())                       # (bin, region) pairs of this rank's slab
            table_bytes = (w["iterations"] + 1) * 243 * 4
            f_sample = F_F5 + F_APPROX + 20
            samples_step = self.nb_local * w["spp"]
            if k_launches:
                # dominant kernel = the tile-major residual pass: draws the sample, picks its region, evaluates f and the region's interpolant,
                # books the moments.  Its own duration (CUDA events around each of its launches inside the timed region), not the step's.
                k_ms_launch = k_ms / k_launches
                samples_launch = samples_step * args.steps / float(k_launches)
                achieved = samples_launch * f_sample / (k_ms_launch * 1e-3) / 1e12
                roof = {"bound": "fp32", "kernel": "cv_tile_samples_kernel<3,5>", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                        "traffic": NCU_TRAFFIC_BYTES.get(self.name), "kernel_ms_per_launch": k_ms_launch, "kernel_launches_per_step": k_launches / float(args.steps),
                        "kernel_share_of_step": (k_ms / args.steps) / ms_dev, "samples_per_launch": samples_launch, "flops_per_residual_sample": f_sample,
                        "smem_gbs_achieved": samples_launch * 972 / (k_ms_launch * 1e-3) / 1e9,
                        "note": "per launch: residual samples x (158 integrand + 846 interpolant + 20) flop, SURVEY.md §8(d); every sample also reads its region's 243 "
                                "coefficients (972 B) from shared memory (smem_gbs_achieved; 148 SMs x 128 B/clk = 37 TB/s)"}
            else:
                achieved = samples_step * f_sample / per_gpu_s / 1e12
                roof = {"bound": "fp32", "kernel": "whole step (sample-major residual pipeline)", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                        "frac": achieved / fp32_peak, "traffic": None, "flops_per_residual_sample": f_sample}
            # the region->bin control-variate pass: regions are marginalised over the non-binned dimensions once (243 -> 9 values), so a
            # (bin, region) pair costs a 2-D patch integral, not the 876 flop of SURVEY.md §8(d)'s 5-D contraction: reported as pairs/s only
            roof.update({"pairs": pairs, "regions_per_bin": pairs / self.nb_local, "pairs_per_s": pairs / per_gpu_s,
                         "hbm_gbs_achieved": (table_bytes + self.nb_local * 4) / per_gpu_s / 1e9, "hbm_gbs_peak": peaks.get("hbm_gbs")})
        elif w["kind"] == "nc":
            bytes_alg = w["iterations"] * 376.0 + (w["iterations"] + 1) * 120.0 + self.nb_local * 4
            achieved = bytes_alg / per_gpu_s / 1e9
            hbm = peaks.get("hbm_gbs") or 6459.3
            roof = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": None, "bytes_per_split": 376,
                    "evals_per_s": w["iterations"] * 20 / per_gpu_s,
                    "note": "376 B per split + 120 B per region + bins in the final pass (SURVEY.md §8d); the step is ~60 dependent refinement rounds of ~10 small launches, "
                            "i.e. launch/latency bound far below the HBM roof"}
        if roof is not None and roof["bound"] == "fp32":
            roof["peak_source"] = f"nominal FP32 (non-tensor) peak 2*128*{ctx.sm_count} SMs*{sm_max:.0f} MHz; MEASURED_PEAKS.json has no FP32 entry (hbm_gbs/bf16 only)"
            roof["peak_measured_fma"] = fma_peak
            roof["frac_of_measured_fma"] = (roof["achieved"] / fma_peak) if fma_peak else None
        cfg = config_of(self.name)
        cfg.update({"rng": RNG_NOTE[w["kind"]] if w["kind"] in RNG_NOTE else None, "parallelism": (f"{self.world} replicas of the whole {self.gres[0]}x{self.gres[1]} grid, one per GPU, own seeds (weak scaling: " + ("the path length of a walk depends on the bin's place in the image, so slabs of a taller grid would not be equal work)" if w["kind"] == "walk" else "the region table does not shard)") if self.replicas
                                    else f"bin-grid slabs x{self.world} ({self.scaling} scaling: global grid {self.gres[0]}x{self.gres[1]})"),
                    "l2": "flushed between timed steps (256 MiB memset, untimed); per-step CUDA events on the library stream"})
        if w["kind"] in ("nc", "cv"):
            cfg["generation"] = "batched top-k refinement (batch = 0); the exact greedy mode (batch = 1) is timed beside it in exact_mode"
        if w["kind"] == "cv" and self.world > 1:
            cfg["table"] = ("rank 0 generates, vb200_regions_broadcast (ncclBroadcast over NVLink) hands the table out" if args.table == "broadcast"
                            else "every rank generates the identical table (no exchange)")
        if philox is not None:
            cfg["philox_value"] = philox
        rec = {"metric": METRIC.get(self.name, f"integrand evals/sec ({w['desc']})"), "value": value, "unit": unit, "n_gpus": self.world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_value, "higher_is_better": True, "scaling": self.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
               "value_source": "burst (agrees with sustained within 3 %)" if agree else "sustained (burst differs by more than 3 %)",
               "burst": {"value": burst, "ms_per_step": ms_dev, "steps": args.steps},
               "sustained": {"value": sus, "ms_per_step": ms_sus, "steps": n_sus, "seconds": ms_sus * n_sus * 1e-3, "sm_mhz_median": sus_clock},
               "e2e": dict({"value": units / (ms_e2e * 1e-3), "unit": unit, "ms_per_step": ms_e2e, "h2d_bytes_per_step": 128, "d2h_bytes_per_step": self.nb_local * 4,
                            "note": "host bins (ordinary pageable numpy memory) through the C ABI: inputs are ~128 B of parameters; the bin slab comes back over PCIe and the "
                                    "reference's '+=' / '=' is applied to the caller's bins on the host, all inside the timed region"}, **rec_extra),
               "gpu_launches": int(round(launches_per_step * args.steps)), "gpu_launches_per_step": launches_per_step, "roofline": roof}
        if exact is not None:
            rec["exact_mode"] = exact
        return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"])
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of back-to-back steps for the sustained figure")
    ap.add_argument("--exact", action="store_true", help="c3: also time the exact greedy mode (seconds)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cv", action="store_true", help="default run: skip the C4 record under 'cv'")
    ap.add_argument("--table", default="replicate", choices=["replicate", "broadcast"],
                    help="c4 at N > 1: every rank generates the region table (default) or rank 0 generates and vb200_regions_broadcast hands it out")
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    primary = args.workload or "c2"

    if args.impl == "reference":
        if rank != 0:
            return 0
        cb, sec, units, done = run_cpu(primary, max(1, args.steps), args.warmup, wall_budget_s=150.0)
        line = {"impl": "reference", "metric": METRIC.get(primary, f"integrand evals/sec ({WORKLOADS[primary]['desc']})"), "value": cb["value"], "unit": cb["unit"], "n_gpus": args.gpus,
                "steps": done, "steps_requested": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "weak" if WORKLOADS[primary]["kind"] in ("mc", "walk") else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_of(primary), "cpu_baseline": cb,
                "note": "CPU reference arm: uses no GPU; steps capped by a 150 s wall budget (each step is a >= 10 s sample, see cpu_baseline.sample)",
                "e2e": {"value": cb["value"], "unit": cb["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from viltrum_b200 import Context
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: viltrum_b200 has no CPU fallback"})); return 1
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    ctx = Context(local_rank)
    if world > 1 and args.table == "broadcast":
        ctx.comm_init_from_torch()           # the library's own NCCL communicator; the rendezvous token travels through torch.distributed
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    fma_peak = None
    try:
        fma_peak = ctx.measure_fp32_peak(5)
    except Exception:
        pass
    time.sleep(0.05)
    sm_max = (sampler.sm_max if sampler and sampler.sm_max else None) or peaks.get("sm_max_mhz") or 1965.0
    fp32_peak = 2 * 128 * ctx.sm_count * sm_max * 1e6 / 1e12             # TFLOP/s nominal: 2 x 128 lanes x SMs x f (SURVEY.md §8d "Peaks")

    line = Bench(primary, args, rank, local_rank, world, ctx, stream, sampler, flush).run(peaks, fp32_peak, fma_peak, sm_max)
    cv = None
    if args.workload is None and not args.no_cv:                          # the CV half of the headline metric rides along in the default run
        cv = Bench("c4", args, rank, local_rank, world, ctx, stream, sampler, flush).run(peaks, fp32_peak, fma_peak, sm_max)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return 0
    if world == 1 and not args.no_cpu_baseline:
        for rec, name in ((line, primary), (cv, "c4")):
            if rec is None:
                continue
            try:
                rec["cpu_baseline"], _, _, _ = run_cpu(name, 1, 0, wall_budget_s=30.0)
            except Exception as ex:      # the baseline is reported, never required
                rec["cpu_baseline"] = {"error": str(ex)}
    else:
        line["cpu_baseline"] = None
    line["clocks"] = clocks
    if cv is not None:
        line["cv"] = cv
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
