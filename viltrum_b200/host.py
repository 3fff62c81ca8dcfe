"""Host-side mirror of the reference's integrator vocabulary for the hot path, over the C ABI.

The reference is header-only C++ (viltrum::integrate(integrator, bins, resolution, f, range), reference
src/integrate.h:72-173); its real drop-in lives in include/viltrum_b200/viltrum.h.  This module re-states the same
names in Python so that the parity tests and bench.py read like the reference's own examples:

    ctx  = Context(0)
    bins = numpy.zeros(1024*1024, numpy.float32)                      # host bins  -> end-to-end path
    integrate(monte_carlo_per_bin_parallel(64, seed=0), bins, [1024, 1024], "shade4_64", range_primary(4), ctx=ctx)

Integrands are the library's built-in synthetic functors, named by string (a Python callable cannot run on the
GPU; C++ users pass their own __device__ functors through the headers).  ``bins`` is a flat float32 buffer in the
reference's tensor layout (dim 0 fastest): a numpy array (staged through the device inside the call) or a CUDA
torch tensor / DevicePtr (used in place).  Write semantics follow the reference integrator by integrator
('+=' vs '=').  Everything runs on the GPU; there is no CPU fallback.
"""
import ctypes
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _capi as C


# ---- ranges (reference src/range.h, src/range-infinite.h) ------------------------------------------------------
@dataclass
class Range:
    min: Sequence[float]
    max: Sequence[float]

    @property
    def dim(self):
        return len(self.min)


@dataclass
class RangeInfinite:
    min: Sequence[float] = ()
    max: Sequence[float] = ()


def range_primary(n):
    """range_primary<N>() — the unit box (reference src/range.h:196-199)"""
    return Range([0.0] * n, [1.0] * n)


def range_primary_infinite():
    """range_primary_infinite<float>() (reference src/range-infinite.h:129-132)"""
    return RangeInfinite()


# ---- device buffers ----------------------------------------------------------------------------------------------
@dataclass
class DevicePtr:
    """A raw device pointer + element count (for callers that manage device memory themselves)."""
    ptr: int
    numel: int


def _buffer(x, dtype=np.float32):
    """-> (address, memory-space flag, keepalive)"""
    if x is None:
        return None, C.HOST, None
    if isinstance(x, DevicePtr):
        return x.ptr, C.DEVICE, x
    if isinstance(x, np.ndarray):
        if x.dtype != dtype or not x.flags["C_CONTIGUOUS"]:
            raise TypeError(f"host buffers must be C-contiguous {np.dtype(dtype).name} arrays (got {x.dtype})")
        return x.ctypes.data, C.HOST, x
    if hasattr(x, "data_ptr") and hasattr(x, "is_cuda"):      # torch tensor, used as plain device memory
        if not x.is_contiguous():
            raise TypeError("tensors must be contiguous")
        return x.data_ptr(), (C.DEVICE if x.is_cuda else C.HOST), x
    raise TypeError(f"unsupported buffer type {type(x)}")


def builtin_names():
    L = C.lib()
    return [L.vb200_builtin_name(i).decode() for i in range(L.vb200_builtin_count())]


class Context:
    """One per process and GPU (vb200_create).  Raises Vb200Error(VB200_ERR_NO_DEVICE) without a CUDA device."""

    def __init__(self, device=0):
        self._L = C.lib()
        h = ctypes.c_void_p()
        rc = self._L.vb200_create(int(device), ctypes.byref(h))
        if rc != 0:
            raise C.Vb200Error(rc, self._L.vb200_last_error(None).decode())
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._L.vb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            raise C.Vb200Error(rc, self._L.vb200_last_error(self._h).decode())

    def synchronize(self):
        self.check(self._L.vb200_synchronize(self._h))

    @property
    def stream(self):
        """cudaStream_t (as int) every call of this context is enqueued on"""
        return self._L.vb200_stream(self._h)

    @property
    def sm_count(self):
        return self._L.vb200_sm_count(self._h)

    @property
    def launch_count(self):
        return self._L.vb200_launch_count(self._h)

    def kernel_timer(self, enable):
        """event pairs around every launch of the residual-sampling kernel of cv_integrate (vb200_kernel_timer)"""
        self.check(self._L.vb200_kernel_timer(self._h, 1 if enable else 0))

    def kernel_timer_read(self):
        """(summed milliseconds, launches) since the last read; synchronises the stream"""
        import ctypes
        ms = ctypes.c_double(0.0); n = ctypes.c_uint64(0)
        self.check(self._L.vb200_kernel_timer_read(self._h, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, int(n.value)

    def host_register(self, array):
        """pin + map a numpy array (vb200_host_register): samplers then write their bins into it directly over PCIe, with no host-side pass"""
        self.check(self._L.vb200_host_register(self._h, array.ctypes.data, array.nbytes))

    def host_unregister(self, array):
        self.check(self._L.vb200_host_unregister(self._h, array.ctypes.data))

    def measure_fp32_peak(self, reps=5):
        """measured dependent-FFMA peak of this GPU in TFLOP/s (vb200_measure_fp32_peak)"""
        out = ctypes.c_double(0.0)
        self.check(self._L.vb200_measure_fp32_peak(self._h, int(reps), ctypes.byref(out)))
        return out.value

    # -- multi-GPU: an optional NCCL communicator owned by the context (one process per GPU) ------------------------------------
    def comm_unique_id(self):
        """128-byte rendezvous token (ncclGetUniqueId): make it on one rank, hand it to the others"""
        buf = (ctypes.c_ubyte * C.COMM_ID_BYTES)()
        self.check(self._L.vb200_comm_unique_id(self._h, buf))
        return bytes(buf)

    def comm_init(self, token, rank, world):
        """ncclCommInitRank on this context's device (collective)"""
        buf = (ctypes.c_ubyte * C.COMM_ID_BYTES).from_buffer_copy(token)
        self.check(self._L.vb200_comm_init(self._h, buf, int(rank), int(world)))

    def comm_init_from_torch(self):
        """convenience for hosts that already run torch.distributed: rank 0 makes the token, the process group's broadcast hands it out"""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        t = torch.zeros(C.COMM_ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            t = torch.frombuffer(bytearray(self.comm_unique_id()), dtype=torch.uint8).clone()
        dev = torch.device("cuda", self.device) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = t.to(dev)
        dist.broadcast(t, 0)
        self.comm_init(bytes(t.cpu().numpy().tobytes()), rank, world)

    def comm_destroy(self):
        self.check(self._L.vb200_comm_destroy(self._h))

    @property
    def comm_rank(self):
        return self._L.vb200_comm_rank(self._h)

    @property
    def comm_size(self):
        return self._L.vb200_comm_size(self._h)

    def regions_broadcast(self, regions, root=0):
        """vb200_regions_broadcast: the root passes its Regions, the other ranks pass None and receive a new table"""
        h = regions._h if regions is not None else ctypes.c_void_p()
        self.check(self._L.vb200_regions_broadcast(self._h, ctypes.byref(h), int(root)))
        return regions if regions is not None else Regions(self, h)

    def integrand(self, name, exact=False):
        if isinstance(name, FubiniIntegrand):      # an adapter descriptor made by vb200_builtin_fubini (always the fast flavour)
            return name.ptr
        p = self._L.vb200_builtin_integrand(name.encode(), 1 if exact else 0)
        if not p:
            raise KeyError(f"unknown built-in integrand '{name}' (have: {builtin_names()})")
        return p

    # -- raw drivers (1:1 with the C ABI) --------------------------------------------------------------------------
    def _mc_params(self, dim, res, rng, spp, seed, flavor, shard, options=0):
        p = C.McParams()
        p.domain = C.make_domain(dim, res, rng.min, rng.max)
        p.shard.begin, p.shard.end = C.shard_pair(shard)
        p.spp, p.seed, p.flavor, p.options = int(spp), int(seed) & 0xFFFFFFFFFFFFFFFF, flavor, int(options)
        return p

    @staticmethod
    def mc_options(rng="xoshiro", lattice24=False):
        """vb200_mc_params.options: rng = 'xoshiro' (default: one xoshiro128++ stream per bin and lane sub-stream, seeded by Philox4x32-10)
        or 'philox' (every draw is Philox4x32-10 of (bin, sample group, call)); lattice24 = True keeps 24 random bits per coordinate
        even inside the bins of a fine grid."""
        if rng not in ("xoshiro", "philox"):
            raise ValueError(f"rng must be 'xoshiro' or 'philox' (got {rng!r})")
        return (C.MC_RNG_PHILOX if rng == "philox" else 0) | (C.MC_LATTICE24 if lattice24 else 0)

    @staticmethod
    def _empty(shard):
        """an explicit empty shard (a rank with no rows): nothing to do — {0,0} in the C ABI means 'whole grid'"""
        return shard is not None and shard[0] == shard[1]

    def mc_per_bin(self, f, bins, res, rng, spp, seed, flavor=C.MC_PER_BIN, shard=None, sum_f=None, sum_f2=None, exact=False,
                   generator="xoshiro", lattice24=False):
        if self._empty(shard):
            return
        b, mem, _k = _buffer(bins)
        s1, m1, _k1 = _buffer(sum_f); s2, m2, _k2 = _buffer(sum_f2)
        p = self._mc_params(len(rng.min), res, rng, spp, seed, flavor, shard, self.mc_options(generator, lattice24))
        self.check(self._L.vb200_mc_per_bin(self._h, self.integrand(f, exact), ctypes.byref(p), b, mem, s1, s2))

    def mc_per_bin_replay(self, f, bins, res, rng, spp, samples, flavor=C.MC_PER_BIN, shard=None, exact=True):
        if self._empty(shard):
            return
        b, mem, _k = _buffer(bins); s, smem, _ks = _buffer(samples)
        p = self._mc_params(len(rng.min), res, rng, spp, 0, flavor, shard)
        self.check(self._L.vb200_mc_per_bin_replay(self._h, self.integrand(f, exact), ctypes.byref(p), s, smem, b, mem))

    def mc_per_bin_inf(self, f, bins, res, rng, spp, seed, shard=None, sum_f=None, sum_f2=None, exact=False, flavor=C.MC_PER_BIN):
        if self._empty(shard):
            return
        b, mem, _k = _buffer(bins)
        s1, m1, _k1 = _buffer(sum_f); s2, m2, _k2 = _buffer(sum_f2)
        p = self._mc_params(len(rng.min), res, rng, spp, seed, flavor, shard)
        self.check(self._L.vb200_mc_per_bin_inf(self._h, self.integrand(f, exact), ctypes.byref(p), b, mem, s1, s2))

    def mc_per_bin_inf_replay(self, f, bins, res, rng, spp, offsets, elems, shard=None, exact=True, flavor=C.MC_PER_BIN):
        if self._empty(shard):
            return
        b, mem, _k = _buffer(bins)
        o, omem, _ko = _buffer(offsets, np.uint64); e, emem, _ke = _buffer(elems)
        p = self._mc_params(len(rng.min), res, rng, spp, 0, flavor, shard)
        self.check(self._L.vb200_mc_per_bin_inf_replay(self._h, self.integrand(f, exact), ctypes.byref(p), o, e, emem, b, mem))

    def monte_carlo(self, f, bins, res, rng, samples, seed, shard=None, exact=False, allreduce=False):
        """allreduce=True: split-sample mode over the context's communicator (VB200_MC_ALLREDUCE) — every rank ends with bins += the TOTAL"""
        if self._empty(shard):
            return
        b, mem, _k = _buffer(bins)
        p = self._mc_params(len(rng.min), res, rng, samples, seed, C.MC_PER_BIN, shard, C.MC_ALLREDUCE if allreduce else 0)
        self.check(self._L.vb200_monte_carlo(self._h, self.integrand(f, exact), ctypes.byref(p), b, mem))

    @staticmethod
    def _fill_mixed(p, mixed):
        """mixed = dict(metric_rest, dimension, bins_weight, size_threshold_bins, size_threshold_rest, error_increase_factor) — error_heuristic_mixed's
        constructor arguments with the reference's defaults (error-heuristic.h:62-68)"""
        m = dict(metric_rest="absolute", dimension=2, bins_weight=1.0, size_threshold_bins=1.0 / 1024.0, size_threshold_rest=1.0 / 16.0, error_increase_factor=1.e4)
        m.update(mixed or {})
        p.mixed.metric_rest, p.mixed.dimension = C.METRICS[m["metric_rest"]], int(m["dimension"])
        p.mixed.bins_weight, p.mixed.size_threshold_bins = float(m["bins_weight"]), float(m["size_threshold_bins"])
        p.mixed.size_threshold_rest, p.mixed.error_increase_factor = float(m["size_threshold_rest"]), float(m["error_increase_factor"])

    def regions_generate_adaptive(self, f, rng, rule, heuristic, metric, iterations, size_weight=1e-5, batch=1, exact=True, mixed=None):
        p = C.AdaptiveParams()
        p.domain = C.make_domain(len(rng.min), [1], rng.min, rng.max)
        p.rule, p.heuristic, p.metric = C.RULES[rule], C.HEURISTICS[heuristic], C.METRICS[metric]
        p.batch, p.size_weight, p.iterations = int(batch), float(size_weight), int(iterations)
        if heuristic == "mixed":
            self._fill_mixed(p, mixed)
        h = ctypes.c_void_p()
        self.check(self._L.vb200_regions_generate_adaptive(self._h, self.integrand(f, exact), ctypes.byref(p), ctypes.byref(h)))
        return Regions(self, h)

    def regions_generate_adaptive_f64(self, f, rng, rule, heuristic, metric, iterations, size_weight=1e-5, exact=True, mixed=None):
        """Range<double,DIM> through the exact greedy generator (vb200_regions_generate_adaptive_f64): a double region table"""
        p = C.AdaptiveParams64()
        p.domain = C.make_domain64(len(rng.min), [1], rng.min, rng.max)
        p.rule, p.heuristic, p.metric = C.RULES[rule], C.HEURISTICS[heuristic], C.METRICS[metric]
        p.batch, p.size_weight, p.iterations = 1, float(size_weight), int(iterations)
        if heuristic == "mixed":
            self._fill_mixed(p, mixed)
        h = ctypes.c_void_p()
        self.check(self._L.vb200_regions_generate_adaptive_f64(self._h, self.integrand64(f, exact), ctypes.byref(p), ctypes.byref(h)))
        return Regions(self, h, f64=True)

    def regions_generate_tolerance(self, f, rng, rule, heuristic, metric, tolerance, size_weight=1e-5, max_regions=0, exact=True):
        p = C.ToleranceParams()
        p.domain = C.make_domain(len(rng.min), [1], rng.min, rng.max)
        p.rule, p.heuristic, p.metric = C.RULES[rule], C.HEURISTICS[heuristic], C.METRICS[metric]
        p.tolerance, p.size_weight, p.max_regions = float(tolerance), float(size_weight), int(max_regions)
        h = ctypes.c_void_p()
        self.check(self._L.vb200_regions_generate_tolerance(self._h, self.integrand(f, exact), ctypes.byref(p), ctypes.byref(h)))
        return Regions(self, h)

    def regions_generate_single(self, f, rng, rule, exact=True):
        d = C.make_domain(len(rng.min), [1], rng.min, rng.max)
        h = ctypes.c_void_p()
        self.check(self._L.vb200_regions_generate_single(self._h, self.integrand(f, exact), ctypes.byref(d), C.rule_id(rule), ctypes.byref(h)))
        return Regions(self, h)

    # -- double precision (Range<double,DIM>): Newton-Cotes region family -------------------------------------------------
    def integrand64(self, name, exact=True):
        p = self._L.vb200_builtin_integrand_f64(name.encode(), 1 if exact else 0)
        if not p:
            raise KeyError(f"unknown double-precision built-in integrand '{name}'")
        return p

    def regions_generate_single_f64(self, f, rng, rule, exact=True):
        d = C.make_domain64(len(rng.min), [1], rng.min, rng.max)
        h = ctypes.c_void_p()
        self.check(self._L.vb200_regions_generate_single_f64(self._h, self.integrand64(f, exact), ctypes.byref(d), C.RULES[rule], ctypes.byref(h)))
        return Regions(self, h, f64=True)

    def regions_upload_f64(self, rule, rmin, rmax, err, errdim, data):
        rmin = np.ascontiguousarray(rmin, np.float64); rmax = np.ascontiguousarray(rmax, np.float64)
        n, dim = rmin.shape
        err = np.ascontiguousarray(err if err is not None else np.zeros(n), np.float64)
        errdim = np.ascontiguousarray(errdim if errdim is not None else np.zeros(n), np.uint32)
        data = np.ascontiguousarray(data, np.float64)
        h = ctypes.c_void_p()
        self.check(self._L.vb200_regions_upload_f64(self._h, dim, C.RULES[rule], n, rmin.ctypes.data, rmax.ctypes.data, err.ctypes.data,
                                                    errdim.ctypes.data, data.ctypes.data, ctypes.byref(h)))
        return Regions(self, h, f64=True)

    def regions_upload(self, rule, rmin, rmax, err, errdim, data):
        rmin = np.ascontiguousarray(rmin, np.float32); rmax = np.ascontiguousarray(rmax, np.float32)
        n, dim = rmin.shape
        err = np.ascontiguousarray(err if err is not None else np.zeros(n), np.float32)
        errdim = np.ascontiguousarray(errdim if errdim is not None else np.zeros(n), np.uint32)
        data = np.ascontiguousarray(data, np.float32)
        h = ctypes.c_void_p()
        self.check(self._L.vb200_regions_upload(self._h, dim, C.RULES[rule], n, rmin.ctypes.data, rmax.ctypes.data, err.ctypes.data,
                                                errdim.ctypes.data, data.ctypes.data, ctypes.byref(h)))
        return Regions(self, h)


def range_split_at(n, rng):
    """range_split_at<N>(range) — reference src/combination/fubini.h:18-49: (first N dimensions, the rest)"""
    if isinstance(rng, RangeInfinite):
        lo = [rng.min[i] if i < len(rng.min) else 0.0 for i in range(n)]
        hi = [rng.max[i] if i < len(rng.max) else 1.0 for i in range(n)]
        return Range(lo, hi), RangeInfinite(list(rng.min[n:]), list(rng.max[n:]))
    return Range(list(rng.min[:n]), list(rng.max[:n])), Range(list(rng.min[n:]), list(rng.max[n:]))


class FubiniIntegrand:
    """function_split_and_integrate_at<N>(f, monte_carlo(mc_samples, seed), range_rest) over a built-in integrand — reference
    src/combination/fubini.h:51-75: the N-dimensional integrand g(x) = vol(rest)/m * sum_s f(x (+) r_s) (vb200_builtin_fubini)."""

    def __init__(self, ctx, name, nfirst, rest, mc_samples, seed):
        lo = np.ascontiguousarray(rest.min, np.float32); hi = np.ascontiguousarray(rest.max, np.float32)
        self._L = ctx._L
        self.name, self.nfirst = name, nfirst
        self.ptr = self._L.vb200_builtin_fubini(name.encode(), int(nfirst), lo.ctypes.data if len(lo) else None, hi.ctypes.data if len(hi) else None,
                                                len(lo), int(mc_samples), int(seed) & 0xFFFFFFFFFFFFFFFF)
        if not self.ptr:
            raise KeyError(f"no built-in Fubini adapter for ('{name}', nfirst={nfirst}) with {len(lo)} rest entries")

    def free(self):
        if getattr(self, "ptr", None):
            self._L.vb200_integrand_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Regions:
    """Device-resident leaf table (vb200_regions)."""

    def __init__(self, ctx, handle, f64=False):
        self.ctx, self._h, self.f64 = ctx, handle, f64

    def __len__(self):
        return int(self.ctx._L.vb200_regions_count(self._h))

    @property
    def dim(self):
        return self.ctx._L.vb200_regions_dim(self._h)

    @property
    def samples(self):
        return self.ctx._L.vb200_regions_samples(self._h)

    def download(self):
        n, d, sd = len(self), self.dim, self.samples
        ft = np.float64 if self.f64 else np.float32
        out = dict(min=np.zeros((n, d), ft), max=np.zeros((n, d), ft), err=np.zeros(n, ft),
                   dim=np.zeros(n, np.uint32), data=np.zeros((n, sd), ft))
        fn = self.ctx._L.vb200_regions_download_f64 if self.f64 else self.ctx._L.vb200_regions_download
        self.ctx.check(fn(self.ctx._h, self._h, out["min"].ctypes.data, out["max"].ctypes.data,
                          out["err"].ctypes.data, out["dim"].ctypes.data, out["data"].ctypes.data))
        return out

    def integrate_bins(self, bins, res, rng, shard=None):
        if Context._empty(shard):
            return
        if self.f64:
            b, mem, _k = _buffer(bins, np.float64)
            d = C.make_domain64(len(rng.min), res, rng.min, rng.max)
            s = C.Shard(); s.begin, s.end = C.shard_pair(shard)
            self.ctx.check(self.ctx._L.vb200_regions_integrate_bins_f64(self.ctx._h, self._h, ctypes.byref(d), ctypes.byref(s), b, mem))
            return
        b, mem, _k = _buffer(bins)
        d = C.make_domain(len(rng.min), res, rng.min, rng.max)
        s = C.Shard(); s.begin, s.end = C.shard_pair(shard)
        self.ctx.check(self.ctx._L.vb200_regions_integrate_bins(self.ctx._h, self._h, ctypes.byref(d), ctypes.byref(s), b, mem))

    def _cv_params(self, res, rng, spp, seed, shard, fixed_alpha=None, rr="uniform", rs="uniform", rs_power=1.0, rs_cutoff=0.0):
        p = C.CvParams()
        p.rr_policy = C.RR_POLICIES[rr]      # rr_uniform_region / rr_integral_region / rr_error_region / rr_pdf_region (region-russian-roulette.h:9-147), "stratified"
        p.rs_policy, p.rs_power, p.rs_cutoff = C.RS_POLICIES[rs], float(rs_power), float(rs_cutoff)      # region-sampling.h:9-135
        p.domain = C.make_domain(len(rng.min), res, rng.min, rng.max)
        p.shard.begin, p.shard.end = C.shard_pair(shard)
        p.spp, p.seed = int(spp), int(seed) & 0xFFFFFFFFFFFFFFFF
        if fixed_alpha is not None:          # cv_fixed_weight(alpha) instead of cv_optimize_weight (weight-strategy.h:7-35)
            p.weight_strategy, p.alpha = C.CV_FIXED_WEIGHT, float(fixed_alpha)
        return p

    def cv_integrate(self, f, bins, res, rng, spp, seed, shard=None, nregions=None, approx=None, exact=False, fixed_alpha=None, rr="uniform",
                     rs="uniform", rs_power=1.0, rs_cutoff=0.0):
        if Context._empty(shard):
            return
        b, mem, _k = _buffer(bins)
        n, _m, _kn = _buffer(nregions, np.uint32); a, _m2, _ka = _buffer(approx)
        p = self._cv_params(res, rng, spp, seed, shard, fixed_alpha, rr, rs, rs_power, rs_cutoff)
        self.ctx.check(self.ctx._L.vb200_cv_integrate(self.ctx._h, self.ctx.integrand(f, exact), self._h, ctypes.byref(p), b, mem, n, a))

    def cv_replay(self, f, bins, res, rng, spp, chosen, samples, shard=None, exact=True, fixed_alpha=None, rr="uniform"):
        if Context._empty(shard):
            return
        b, mem, _k = _buffer(bins)
        c, cmem, _kc = _buffer(chosen, np.uint32); s, _sm, _ks = _buffer(samples)
        p = self._cv_params(res, rng, spp, 0, shard, fixed_alpha, rr)
        self.ctx.check(self.ctx._L.vb200_cv_replay(self.ctx._h, self.ctx.integrand(f, exact), self._h, ctypes.byref(p), c, s, cmem, b, mem))

    def free(self):
        if self._h:
            self.ctx._L.vb200_regions_free(self._h)      # safe after Context.close(): the library orphans outstanding tables in vb200_destroy
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---- integrators: same factory names and argument meaning as the reference ---------------------------------------
@dataclass
class MonteCarlo:
    """monte_carlo(samples, seed) — reference src/monte-carlo/monte-carlo.h:88-95"""
    samples: int
    seed: int = 0

    def integrate(self, ctx, bins, res, f, rng, shard=None, **kw):
        ctx.monte_carlo(f, bins, res, rng, self.samples, self.seed, shard=shard, **kw)


@dataclass
class MonteCarloPerBinParallel:
    """monte_carlo_per_bin_parallel(spp, seed) — reference src/monte-carlo/monte-carlo-per-bin-parallel.h:104-111 ('+=')"""
    spp: int
    seed: int = 0

    def integrate(self, ctx, bins, res, f, rng, shard=None, **kw):
        if isinstance(rng, RangeInfinite):
            ctx.mc_per_bin_inf(f, bins, res, rng, self.spp, self.seed, shard=shard, **kw)
        else:
            ctx.mc_per_bin(f, bins, res, rng, self.spp, self.seed, C.MC_PER_BIN, shard=shard, **kw)


@dataclass
class IntegratorPerBinParallel:
    """integrator_per_bin_parallel(inner) — reference src/integrator-per-bin-parallel.h:37-38 ('=')"""
    inner: MonteCarlo

    def integrate(self, ctx, bins, res, f, rng, shard=None, **kw):
        if not isinstance(self.inner, MonteCarlo):
            raise NotImplementedError("integrator_per_bin_parallel: only monte_carlo(...) inner integrators are on the hot path")
        if isinstance(rng, RangeInfinite):
            ctx.mc_per_bin_inf(f, bins, res, rng, self.inner.samples, self.inner.seed, shard=shard, flavor=C.PER_BIN_MC, **kw)
        else:
            ctx.mc_per_bin(f, bins, res, rng, self.inner.samples, self.inner.seed, C.PER_BIN_MC, shard=shard, **kw)


@dataclass
class Nested:
    high: str
    low: str

    @property
    def name(self):
        return f"{self.high}_{self.low}"


@dataclass
class ErrorMetric:
    kind: str


@dataclass
class ErrorHeuristic:
    kind: str
    metric: ErrorMetric
    size_weight: float = 1e-5


def nested(high, low):
    """nested(high, low) — reference src/nested/nested.h:42-45; rules are named 'trapezoidal' | 'simpson' | 'boole'"""
    return Nested(high, low)


def error_metric_absolute():
    return ErrorMetric("absolute")


def error_metric_relative():
    return ErrorMetric("relative")


def error_heuristic_default(metric):
    return ErrorHeuristic("default", metric)


def error_heuristic_size(metric, size_weight=1e-5):
    return ErrorHeuristic("size", metric, size_weight)


def error_heuristic_mixed(metric_bins, metric_rest, dimension=2, bins_weight=1.0, size_weight=1.e-3, size_threshold_bins=1.0 / 1024.0,
                          size_threshold_rest=1.0 / 16.0, error_increase_factor=1.e4):
    """error_heuristic_mixed(...) — reference src/nested/error-heuristic.h:49-98, same argument order and defaults"""
    h = ErrorHeuristic("mixed", metric_bins, size_weight)
    h.mixed = dict(metric_rest=metric_rest.kind, dimension=dimension, bins_weight=bins_weight, size_threshold_bins=size_threshold_bins,
                   size_threshold_rest=size_threshold_rest, error_increase_factor=error_increase_factor)
    return h


@dataclass
class IntegratorNewtonCotes:
    """integrator_newton_cotes(rule) — reference src/newton-cotes/newton-cotes.h:11-14 ('+=')"""
    rule: str

    def integrate(self, ctx, bins, res, f, rng, shard=None, exact=True, **kw):
        regs = ctx.regions_generate_single(f, rng, self.rule, exact=exact)
        try:
            regs.integrate_bins(bins, res, rng, shard=shard)
        finally:
            regs.free()


@dataclass
class IntegratorAdaptiveIterations:
    """integrator_adaptive_iterations(nested_rule, error_heuristic, iterations) — reference
    src/nested/integrator-adaptive-iterations.h:12-15 ('+=').  batch=1 reproduces the reference's greedy order."""
    rule: Nested
    heuristic: ErrorHeuristic
    iterations: int
    batch: int = 1
    last_regions: Optional[Regions] = field(default=None, repr=False)

    def generate(self, ctx, f, rng, exact=True):
        return ctx.regions_generate_adaptive(f, rng, self.rule.name, self.heuristic.kind, self.heuristic.metric.kind,
                                             self.iterations, self.heuristic.size_weight, self.batch, exact=exact, mixed=getattr(self.heuristic, "mixed", None))

    def integrate(self, ctx, bins, res, f, rng, shard=None, exact=True, logger=None, **kw):
        regs = self.generate(ctx, f, rng, exact=exact)
        if logger is not None:
            logger.log(regs)          # the reference hands the region list to Logger::log (integrator-region-based.h:19)
        try:
            regs.integrate_bins(bins, res, rng, shard=shard)
        finally:
            if logger is None:
                regs.free()


@dataclass
class IntegratorAdaptiveTolerance:
    """integrator_adaptive_tolerance(nested_rule, error_heuristic, tolerance) — reference src/nested/integrator-adaptive-tolerance.h:41-59
    ('+=').  Leaves come out in the reference's depth-first order, so the bins match bit for bit with an exact integrand."""
    rule: Nested
    heuristic: ErrorHeuristic
    tolerance: float = 1e-3
    max_regions: int = 0

    def generate(self, ctx, f, rng, exact=True):
        return ctx.regions_generate_tolerance(f, rng, self.rule.name, self.heuristic.kind, self.heuristic.metric.kind, self.tolerance,
                                              self.heuristic.size_weight, self.max_regions, exact=exact)

    def integrate(self, ctx, bins, res, f, rng, shard=None, exact=True, logger=None, **kw):
        regs = self.generate(ctx, f, rng, exact=exact)
        if logger is not None:
            logger.log(regs)
        try:
            regs.integrate_bins(bins, res, rng, shard=shard)
        finally:
            if logger is None:
                regs.free()


@dataclass
class CvFixedWeight:
    """cv_fixed_weight(alpha) — reference src/control-variates/weight-strategy.h:7-35"""
    alpha: float = 1.0


def cv_fixed_weight(alpha=1.0):
    return CvFixedWeight(alpha)


def cv_optimize_weight():
    """cv_optimize_weight() — reference src/control-variates/weight-strategy.h:40-110 (the default)"""
    return None


def rr_uniform_region():
    """rr_uniform_region() — reference src/control-variates/region-russian-roulette.h:9-28 (the crespo2021 preset)"""
    return "uniform"


def rr_integral_region():
    """rr_integral_region() — reference src/control-variates/region-russian-roulette.h:30-67"""
    return "integral"


def rr_error_region():
    """rr_error_region() — reference src/control-variates/region-russian-roulette.h:69-106"""
    return "error"


def rr_pdf_region():
    """rr_pdf_region() — reference src/control-variates/region-russian-roulette.h:108-147 (factor_prob = 0.01)"""
    return "pdf"


def region_stratification_uniform():
    """region_stratification_uniform() — reference src/control-variates/region-stratification.h:9-25 (the Optimized integrator's allocation)"""
    return "stratified"


@dataclass
class RegionSampling:
    """region_sampling_uniform / _importance / _mis(power, cutoff) / _russian_roulette — reference src/control-variates/region-sampling.h:9-135"""
    kind: str = "uniform"
    power: float = 1.0
    cutoff: float = 0.0


def region_sampling_uniform():
    return RegionSampling("uniform")


def region_sampling_importance():
    return RegionSampling("importance")


def region_sampling_mis(power=1.0, cutoff=0.0):
    return RegionSampling("mis", power, cutoff)


def region_sampling_russian_roulette():
    return RegionSampling("russian_roulette")


@dataclass
class IntegratorCrespo2021:
    """integrator_crespo2021(iterations, spp, seed) — reference src/control-variates/integrator-crespo2021.h:7-22 ('=').
    cv=cv_fixed_weight(alpha) gives integrator_adaptive_variance_reduction_parallel(..., rr_uniform_region(), cv_fixed_weight(alpha), ...)."""
    iterations: int
    spp: int
    seed: int = 0
    batch: int = 1
    cv: Optional[CvFixedWeight] = None
    rr: str = "uniform"
    rs: Optional["RegionSampling"] = None

    def integrate(self, ctx, bins, res, f, rng, shard=None, exact=False, logger=None, **kw):
        gen = IntegratorAdaptiveIterations(nested("simpson", "trapezoidal"), error_heuristic_size(error_metric_relative(), 1e-5),
                                           self.iterations, self.batch)
        regs = gen.generate(ctx, f, rng, exact=True)
        if logger is not None:
            logger.log(regs)
        try:
            rs = self.rs or RegionSampling()
            regs.cv_integrate(f, bins, res, rng, self.spp, self.seed, shard=shard, exact=exact,
                              fixed_alpha=self.cv.alpha if self.cv is not None else None, rr=self.rr, rs=rs.kind, rs_power=rs.power, rs_cutoff=rs.cutoff, **kw)
        finally:
            if logger is None:
                regs.free()


@dataclass
class IntegratorFubini:
    """integrator_fubini<N>(first, monte_carlo(m, seed)) — reference src/combination/fubini.h:78-101: `first` integrates the first N
    dimensions of g(x) = the Monte-Carlo estimate of the integral of f(x, .) over the rest (finite or infinite)."""
    nfirst: int
    first: object
    rest: MonteCarlo

    def integrate(self, ctx, bins, res, f, rng, shard=None, **kw):
        if len(res) > self.nfirst:
            raise ValueError("Fubini does not work with that many dimensions on bin resolution")      # fubini.h:88
        first_rng, rest_rng = range_split_at(self.nfirst, rng)
        g = FubiniIntegrand(ctx, f, self.nfirst, rest_rng, self.rest.samples, self.rest.seed)
        try:
            kw.pop("exact", None)
            self.first.integrate(ctx, bins, res, g, first_rng, shard=shard, **kw)
        finally:
            g.free()


@dataclass
class IntegratorCrespo2021Infinite:
    """integrator_crespo2021_infinite<N>(iterations, mc_samples, spp, seed) — reference src/control-variates/integrator-crespo2021.h:24-44:
    the region table is generated over the first N dimensions from g estimated with monte_carlo(mc_samples, 2*seed+1)
    (regions_generator_fubini<N>, regions-generator-fubini.h:7-28); the residual pass evaluates f with ONE sample of the rest per
    residual sample (monte_carlo_per_bin(rng,1), regions-integrator-parallel-variance-reduction.h:69).  '='."""
    nfirst: int
    iterations: int
    mc_samples: int
    spp: int
    seed: int = 0
    batch: int = 1

    def integrate(self, ctx, bins, res, f, rng, shard=None, logger=None, **kw):
        first_rng, rest_rng = range_split_at(self.nfirst, rng)
        g_gen = FubiniIntegrand(ctx, f, self.nfirst, rest_rng, self.mc_samples, 2 * self.seed + 1)
        g_res = FubiniIntegrand(ctx, f, self.nfirst, rest_rng, 1, self.seed + 0x9E3779B97F4A7C15)
        gen = IntegratorAdaptiveIterations(nested("simpson", "trapezoidal"), error_heuristic_size(error_metric_relative(), 1e-5),
                                           self.iterations, self.batch)
        regs = None
        try:
            regs = gen.generate(ctx, g_gen, first_rng)
            if logger is not None:
                logger.log(regs)
            kw.pop("exact", None)
            regs.cv_integrate(g_res, bins, res, first_rng, self.spp, self.seed, shard=shard, **kw)
        finally:
            if regs is not None and logger is None:
                regs.free()
            g_gen.free(); g_res.free()


@dataclass
class IntegratorFubiniVarianceReductionOptimized(IntegratorCrespo2021Infinite):
    """integrator_adaptive_fubini_variance_reduction_parallel_optimized<N>(nested(simpson,trapezoidal), error_heuristic_size(relative), iterations,
    mc_samples, region_stratification_uniform(), cv, region_sampling_uniform(), spp, seed) — reference
    src/control-variates/integrator-adaptive-fubini-variance-reduction-optimized.h:9-23: the crespo2021_infinite pipeline with the stratified
    allocation of RegionsIntegratorParallelVarianceReductionOptimized in place of the per-sample Russian roulette."""

    def integrate(self, ctx, bins, res, f, rng, shard=None, logger=None, **kw):
        kw = dict(kw); kw["rr"] = "stratified"
        super().integrate(ctx, bins, res, f, rng, shard=shard, logger=logger, **kw)


def integrator_adaptive_fubini_variance_reduction_parallel_optimized(nfirst, iterations, mc_samples, spp, seed=0, batch=1):
    return IntegratorFubiniVarianceReductionOptimized(nfirst, iterations, mc_samples, spp, seed, batch)


def integrator_fubini(nfirst, first, rest):
    return IntegratorFubini(nfirst, first, rest)


def integrator_crespo2021_infinite(nfirst, iterations, mc_samples, spp, seed=0, batch=1):
    return IntegratorCrespo2021Infinite(nfirst, iterations, mc_samples, spp, seed, batch)


def monte_carlo(samples, seed=0):
    return MonteCarlo(samples, seed)


def monte_carlo_per_bin_parallel(spp, seed=0):
    return MonteCarloPerBinParallel(spp, seed)


def integrator_per_bin_parallel(inner):
    return IntegratorPerBinParallel(inner)


def steps(n, rule):
    """steps<N>(rule) — reference src/newton-cotes/rules.h:386-389: N pieces of `rule` per dimension (fixed-rule integration only)"""
    return f"steps{int(n)}_{rule}"


def integrator_newton_cotes(rule):
    return IntegratorNewtonCotes(rule)


def integrator_adaptive_iterations(rule, heuristic=None, iterations=None, batch=1):
    if iterations is None:            # integrator_adaptive_iterations(rule, iterations) overload (integrator-adaptive-iterations.h:22-25)
        heuristic, iterations = error_heuristic_default(error_metric_absolute()), heuristic
    return IntegratorAdaptiveIterations(rule, heuristic, iterations, batch)


def integrator_adaptive_tolerance(rule, heuristic=None, tolerance=1e-3, max_regions=0):
    if not isinstance(heuristic, ErrorHeuristic):        # integrator_adaptive_tolerance(rule, tolerance) overload (integrator-adaptive-tolerance.h:46-54)
        if heuristic is not None:
            tolerance = heuristic
        heuristic = error_heuristic_default(error_metric_absolute())
    return IntegratorAdaptiveTolerance(rule, heuristic, tolerance, max_regions)


def integrator_crespo2021(iterations, spp, seed=0, batch=1, cv=None):
    return IntegratorCrespo2021(iterations, spp, seed, batch, cv)


def integrator_adaptive_variance_reduction_parallel(rule, heuristic, iterations, rr, cv, spp, seed=0, batch=1, rs=None):
    """integrator_adaptive_variance_reduction_parallel(nested(simpson,trapezoidal), error_heuristic_size(relative), iterations, RR, CV,
    region_sampling_uniform(), spp, seed) — reference src/control-variates/integrator-adaptive-variance-reduction.h: the crespo2021 rule /
    heuristic pair with RR = rr_uniform_region() | rr_integral_region() | rr_error_region() | rr_pdf_region() and CV = cv_optimize_weight() | cv_fixed_weight(a)."""
    if getattr(rule, "name", rule) != "simpson_trapezoidal":
        raise ValueError("the device control-variate path is built for nested(simpson, trapezoidal)")
    if not (isinstance(heuristic, ErrorHeuristic) and heuristic.kind == "size" and heuristic.metric.kind == "relative"):
        raise ValueError("the device control-variate path is built for error_heuristic_size(error_metric_relative())")
    return IntegratorCrespo2021(iterations, spp, seed, batch, cv, rr, rs)


# ---- multi-GPU partitioning (one process per GPU) ---------------------------------------------------
def shard_for_rank(resolution, rank, world):
    """Bin-grid slab of `rank`: a contiguous range [begin, end) of linear bin indices (tensor order) made of whole rows of the
    LAST bin dimension, so that every rank owns one contiguous block of the flat bin array.  Ranks beyond the number of rows
    get an empty shard (begin == end; the drivers pass it to the C ABI as VB200_SHARD_EMPTY, {0,0} would mean 'whole grid')."""
    resolution = [int(r) for r in resolution]
    rows = resolution[-1]
    stride = 1
    for r in resolution[:-1]:
        stride *= r
    lo = rows * rank // world
    hi = rows * (rank + 1) // world
    return lo * stride, hi * stride


def sample_shard_for_rank(samples, rank, world):
    """Sample-index range of `rank` for the split-sample mode of monte_carlo (few bins, many samples): every rank draws its own
    range of the global sample counter and the partial grids are summed (the one allreduce of the design)."""
    return int(samples) * rank // world, int(samples) * (rank + 1) // world


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def integrate(integrator, bins, resolution, f, rng, ctx=None, **kw):
    """viltrum::integrate(integrator, bins, resolution, f, range) — reference src/integrate.h:72-103.
    ``bins`` is flat (tensor order); pass ``resolution=None`` for the std::vector overload (1-D, res = len(bins),
    integrate.h:132-137)."""
    if resolution is None:
        resolution = [bins.numel() if hasattr(bins, "numel") and callable(bins.numel) else len(bins)]
    ctx = ctx or default_context()
    integrator.integrate(ctx, bins, list(resolution), f, rng, **kw)
    return bins
