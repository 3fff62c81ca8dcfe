#!/usr/bin/env python
"""bench.py — headline benchmark of the per-bin integration hot path (BASELINE.json metric: integrand evals/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c2b|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input: one call of
monte_carlo_per_bin_parallel(64, seed) over a 1024x1024-bin grid of the 4-D shade4<64> integrand
(BASELINE.json configs[1]; SURVEY.md §8d "C2") = 67.1 M integrand evaluations per GPU.  N GPUs shard the bin grid
(weak scaling: every rank owns a 1024x1024 slab of a 1024 x 1024N grid, Philox counters keyed by the global bin index,
no data-path collective — SURVEY.md §8e).

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (bins stay in HBM), `e2e` = the same call with
HOST bins through the C ABI (device->host copy of the 4 MiB bin slab and the host-side '+=' inside the timed region).
`roofline` is the FP32 (non-tensor) roofline SURVEY.md §8(d) names for this kernel (bound "fp32": this path has no tensor-core work
and moves 4 B per 64 evaluations, so neither of the contract's "hbm"/"tensor" bounds describes it; the achieved HBM GB/s is reported
beside it); `cpu_baseline` is the
reference's CPU path timed on this box's host cores (rank 0, N=1, bounded sample).
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

WORKLOADS = {
    # name: (description, integrand, res per GPU, spp, flavor, flops per eval, infinite?)
    "c2": ("per-bin stratified MC, 1024x1024 bins, 64 spp, shade4<64> 4D, fp32 (BASELINE configs[1])", "shade4_64", [1024, 1024], 64, "mc_per_bin_parallel", 155, False),
    "c2b": ("integrator_per_bin_parallel(monte_carlo(64)), 1024x1024 bins, shade4<64>", "shade4_64", [1024, 1024], 64, "per_bin_parallel_mc", 155, False),
    "c2k16": ("per-bin stratified MC, 1024x1024 bins, 64 spp, shade4<16>", "shade4_16", [1024, 1024], 64, "mc_per_bin_parallel", 59, False),
    "c5": ("range_infinite random walk with Russian roulette, 2048x2048 bins, 256 spp (BASELINE configs[4])", "walk", [2048, 2048], 256, "mc_per_bin_parallel_inf", None, True),
}
METRIC = "integrand evals/sec (per-bin MC, 1024x1024 bins x 64 spp, shade4<64>)"
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the ncu --set full capture summarised in
# profiles/ncu_c2_mc_per_bin_r1c.txt / profiles/ncu_c5_walk_window_r1d.txt (algorithmic bytes: 4 B per bin)
NCU_TRAFFIC_BYTES = {"c2": 4231936, "c5": 16922112}
RNG_NOTE = {"mc": "Philox4x32-10, 5 calls per group of 8 samples: 24-bit fields for the free dimensions, 16-bit fields inside a bin of a >=256-bin axis (every generated bit is used)",
            "walk": "Philox4x32-10, counter (bin, sample, block): one block of four 24-bit elements per lane and loop iteration, first-round products cached (18 multiplies per block)"}


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region through NVML (same data as the nvidia-smi
    clocks line of B200_PROFILING.md, but at ~1 kHz so that a millisecond-scale timed region still gets samples)."""

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.stop_flag, self.err = index, [], 0, False, None
        self.sm_max = None

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


def run_cpu(workload, steps, warmup, sample_rows=None):
    """The reference's own CPU implementation of the path on this box's host cores: oracle/_ref (the unmodified reference)
    when it was built, else the oracle port.  Bounded sample: a slab of the workload's bin grid, all host threads
    (thread-pool driver slabbing the last bin dimension)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    desc, integ, res, spp, path, flops, inf = WORKLOADS[workload]
    kind = "reference" if pyoracle.available("reference") or os.path.isdir(pyoracle.REFERENCE_ROOT) else "port"
    O = pyoracle.load(kind)
    T = cpu_threads()
    rows = sample_rows or max(T, min(res[1], 16 * T if inf else 48 * T))
    rows = min(rows, res[1])
    sres = [res[0], rows]
    rmin, rmax = ((), ()) if inf else ([0.0] * O.dim(integ), [1.0] * O.dim(integ))
    units = sres[0] * sres[1] * spp
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        bins = O.mt_per_bin(path, integ, sres, spp, i, T, rmin, rmax)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    assert np.all(np.isfinite(bins))
    sec = float(np.mean(times))
    return dict(value=units / sec, unit="evals/s", cores=T, kind=kind,
                sample=f"{sres[0]}x{sres[1]}-bin slab of the {res[0]}x{res[1]} grid, {spp} spp ({units/1e6:.1f} M evals per step), {T} threads, mean of {steps} steps"), sec, units


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    desc, integ, res, spp, path, flops, inf = WORKLOADS[args.workload]
    metric = METRIC if args.workload == "c2" else f"integrand evals/sec ({desc})"

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
        cb, sec, units = run_cpu(args.workload, steps, warmup)
        line = {"impl": "reference", "metric": metric, "value": cb["value"], "unit": "evals/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
                "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": desc, "integrand": integ, "bins_per_step": [res[0], int(units // (res[0] * spp))], "spp": spp,
                           "note": "CPU reference arm: bounded slab of the workload per step; uses no GPU"},
                "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from viltrum_b200 import Context, Range, RangeInfinite, _capi
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: viltrum_b200 has no CPU fallback"})); return 1
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    ctx = Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    nb_local = res[0] * res[1]
    # weak scaling: the global grid is res[0] x (res[1]*world); this rank owns rows [rank*res[1], (rank+1)*res[1])
    gres = [res[0], res[1] * world]
    from viltrum_b200 import shard_for_rank
    shard = shard_for_rank(gres, rank, world)
    assert shard == (rank * nb_local, (rank + 1) * nb_local)
    rng = RangeInfinite() if inf else Range([0.0] * 4, [1.0] * 4)
    d_bins = torch.zeros(gres[0] * gres[1], dtype=torch.float32, device="cuda")
    h_bins = np.zeros(gres[0] * gres[1], np.float32)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    def step_device(seed):
        if inf:
            ctx.mc_per_bin_inf(integ, d_bins, gres, rng, spp, seed, shard=shard)
        else:
            ctx.mc_per_bin(integ, d_bins, gres, rng, spp, seed, _capi.MC_PER_BIN if path == "mc_per_bin_parallel" else _capi.PER_BIN_MC, shard=shard)

    def step_host(seed):
        if inf:
            ctx.mc_per_bin_inf(integ, h_bins, gres, rng, spp, seed, shard=shard)
        else:
            ctx.mc_per_bin(integ, h_bins, gres, rng, spp, seed, _capi.MC_PER_BIN if path == "mc_per_bin_parallel" else _capi.PER_BIN_MC, shard=shard)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        with torch.cuda.stream(stream):
            for i in range(warmup):
                fn(i)
            barrier()
            evs = []
            for i in range(steps):
                flush.zero_()                                     # L2 flush between timed iterations (untimed)
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream); fn(1000 + i); e1.record(stream)
                evs.append((e0, e1))
            barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs) / steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count
    ms_dev = timed(step_device, args.steps, args.warmup)
    launches = ctx.launch_count - l0 - args.warmup
    ms_e2e = timed(step_host, args.steps, max(3, args.warmup))               # caller's bins in ordinary pageable memory: staged stores + pooled host '+='
    ctx.host_register(h_bins)                                                # optional: caller's bins pinned + mapped once, outside the timed loop
    ms_e2e_pinned = timed(step_host, args.steps, max(3, args.warmup))        # the same call: the kernel applies '+=' to the host bins in place over PCIe
    ctx.host_unregister(h_bins)
    clocks = sampler.stop() if rank == 0 else None
    units = nb_local * spp * world
    value = units / (ms_dev * 1e-3)
    e2e = units / (ms_e2e * 1e-3)
    e2e_pinned = units / (ms_e2e_pinned * 1e-3)
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
    fp32_peak = 2 * 128 * ctx.sm_count * sm_max * 1e6 / 1e12             # TFLOP/s nominal: 2 x 128 lanes x SMs x f (SURVEY.md §8d "Peaks")
    roof = None
    fma_peak = None
    try:
        fma_peak = ctx.measure_fp32_peak(5)
    except Exception:
        pass
    if flops:
        achieved = (nb_local * spp / (ms_dev * 1e-3)) * flops / 1e12 if world == 1 else value / world * flops / 1e12
        roof = {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak, "traffic": NCU_TRAFFIC_BYTES.get(args.workload),
                "peak_source": f"nominal FP32 (non-tensor) peak 2*128*{ctx.sm_count} SMs*{sm_max:.0f} MHz; MEASURED_PEAKS.json has no FP32 entry (hbm_gbs/bf16 only)",
                "peak_measured_fma": fma_peak, "frac_of_measured_fma": (achieved / fma_peak) if fma_peak else None, "flops_per_eval": flops, "hbm_gbs_achieved": nb_local * 4 / (ms_dev * 1e-3) / 1e9, "hbm_gbs_peak": peaks.get("hbm_gbs")}
    cb = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cb, _, _ = run_cpu(args.workload, 2, 1)
        except Exception as ex:      # the baseline is reported, never required
            cb = {"error": str(ex)}
    line = {"metric": metric, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "integrand": integ, "bins_per_gpu": res, "spp": spp, "rng": RNG_NOTE["walk" if args.workload == "c5" else "mc"], "parallelism": f"bin-grid slabs x{world}",
                       "l2": "flushed between timed steps (256 MiB memset, untimed); per-step CUDA events on the library stream"},
            "e2e": {"value": e2e, "unit": "evals/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": 128, "d2h_bytes_per_step": nb_local * 4,
                    "pinned_value": e2e_pinned, "pinned_ms_per_step": ms_e2e_pinned,
                    "note": "host bins (ordinary pageable numpy memory) through vb200_mc_per_bin: one launch stores the bin estimates over PCIe into a pinned staging buffer "
                            "with per-chunk completion flags while a small pool of host threads applies the reference's '+=' to the caller's bins; inputs are ~128 B of "
                            "parameters.  pinned_value: the same call after vb200_host_register(bins) — the kernel reads and writes the caller's bins in place over PCIe "
                            "(faster on a single-GPU host, slower where several ranks share the host's PCIe read path)"},
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cb, "clocks": clocks}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
