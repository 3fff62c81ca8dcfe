"""TEST INFRASTRUCTURE — not product code.

ctypes loader for the two CPU checkers (see oracle/oracle_api.h):

  * ``load("port")``       -> oracle/liboracle.so            plain restatement (oracle/oracle.cpp)
  * ``load("reference")``  -> oracle/_ref/libviltrum_ref.so  the unmodified reference
  * ``load("reference-mt")`` -> oracle/_ref/libviltrum_ref_mt.so  the same, with a std::thread back end under the reference's
    std::for_each(par_unseq, ...) loops (upstream: TBB) — bench.py's multi-threaded CPU baseline, never used by the parity tests

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
this module.  Nothing under viltrum_b200/ does.
"""
import ctypes
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libviltrum_ref.so")
REF_MT_SO = os.path.join(HERE, "_ref", "libviltrum_ref_mt.so")      # timing baseline only (bench.py): see oracle/pstl_threads/execution
_PATHS = {"port": PORT_SO, "reference": REF_SO, "reference-mt": REF_MT_SO}
REFERENCE_ROOT = "/root/reference"

RULE_SAMPLES = {"trapezoidal": 2, "simpson": 3, "boole": 5, "simpson_trapezoidal": 3, "boole_simpson": 5}


def build(kind="port"):
    """(Re)build a checker with oracle/Makefile.  'reference' needs /root/reference (authoring container only)."""
    target = {"port": "port", "reference": "ref", "reference-mt": "ref"}[kind]
    if kind != "port" and not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError("/root/reference is absent: the reference checker can only be built in the authoring container")
    subprocess.run(["make", "-s", "-j8", "-C", HERE, target], check=True)


def available(kind):
    return os.path.exists(_PATHS[kind])


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


class Oracle:
    """Thin numpy front end over the vo_* C ABI.  Bins are flat float32 arrays in viltrum::tensor order
    (dim 0 fastest); ``res`` is the per-dimension bin count."""

    def __init__(self, path):
        self.lib = ctypes.CDLL(path)
        self.lib.vo_kind.restype = ctypes.c_char_p
        self.kind = self.lib.vo_kind().decode()
        for name in ("vo_mc_per_bin_parallel", "vo_per_bin_parallel_mc", "vo_monte_carlo", "vo_mc_per_bin_parallel_inf",
                     "vo_per_bin_parallel_mc_inf", "vo_monte_carlo_inf", "vo_newton_cotes", "vo_adaptive_iterations", "vo_crespo2021", "vo_mt_per_bin", "vo_integrand_dim"):
            getattr(self.lib, name).restype = ctypes.c_int

    def dim(self, integrand):
        return self.lib.vo_integrand_dim(integrand.encode())

    def set_mixed(self, dimension=2, bins_weight=1.0, size_threshold_bins=1.0 / 1024.0, size_threshold_rest=1.0 / 16.0, error_increase_factor=1.e4):
        """constructor arguments of error_heuristic_mixed for the next 'mixed_<bins>_<rest>' heuristic (size_weight travels with the call)"""
        self.lib.vo_set_mixed.restype = None
        self.lib.vo_set_mixed(int(dimension), ctypes.c_double(bins_weight), ctypes.c_double(size_threshold_bins), ctypes.c_double(size_threshold_rest),
                              ctypes.c_double(error_increase_factor))

    def set_threads(self, n):
        """threads behind the reference's par_unseq loops (reference-mt build; 1 elsewhere)"""
        if not hasattr(self.lib, "vo_set_threads"):
            return 1
        self.lib.vo_set_threads.restype = ctypes.c_int
        return self.lib.vo_set_threads(int(n))

    def phase_times(self):
        """(seconds until the last region-based call logged its region list = generation, seconds of the whole call); reference builds only"""
        if not hasattr(self.lib, "vo_phase_times"):
            return None
        t = (ctypes.c_double * 2)()
        self.lib.vo_phase_times.restype = None
        self.lib.vo_phase_times(t)
        return float(t[0]), float(t[1])

    @staticmethod
    def _setup(res, rmin, rmax):
        res = np.ascontiguousarray(np.asarray(res, dtype=np.uint64))
        return res, _f32(rmin), _f32(rmax), int(np.prod(res))

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed in {self.kind} oracle: rc={rc}")

    def _per_bin(self, fn, integrand, res, rmin, rmax, spp, seed, bins, record):
        """record = True: (bins, samples, sum f, sum f^2); record = "moments": the same without the sample points (samples = None)"""
        res, rmin, rmax, nb = self._setup(res, rmin, rmax)
        d = self.dim(integrand)
        bins = np.zeros(nb, np.float32) if bins is None else np.ascontiguousarray(bins, dtype=np.float32).copy()
        samples = np.zeros((nb, spp, d), np.float32) if record is True else None
        s1 = np.zeros(nb, np.float64) if record else None
        s2 = np.zeros(nb, np.float64) if record else None
        rc = fn(integrand.encode(), len(res), _p(res), _p(rmin), _p(rmax), ctypes.c_uint64(spp), ctypes.c_uint64(seed),
                _p(bins), _p(samples), _p(s1), _p(s2))
        self._check(rc, fn.__name__)
        return (bins, samples, s1, s2) if record else bins

    def mc_per_bin_parallel(self, integrand, res, rmin, rmax, spp, seed, bins=None, record=False):
        return self._per_bin(self.lib.vo_mc_per_bin_parallel, integrand, res, rmin, rmax, spp, seed, bins, record)

    def per_bin_parallel_mc(self, integrand, res, rmin, rmax, spp, seed, bins=None, record=False):
        return self._per_bin(self.lib.vo_per_bin_parallel_mc, integrand, res, rmin, rmax, spp, seed, bins, record)

    def monte_carlo(self, integrand, res, rmin, rmax, samples, seed, bins=None, record=False):
        res, rmin, rmax, nb = self._setup(res, rmin, rmax)
        d = self.dim(integrand)
        bins = np.zeros(nb, np.float32) if bins is None else np.ascontiguousarray(bins, dtype=np.float32).copy()
        rec = np.zeros((samples, d), np.float32) if record else None
        rc = self.lib.vo_monte_carlo(integrand.encode(), len(res), _p(res), _p(rmin), _p(rmax), ctypes.c_uint64(samples),
                                     ctypes.c_uint64(seed), _p(bins), _p(rec))
        self._check(rc, "vo_monte_carlo")
        return (bins, rec) if record else bins

    def mc_per_bin_parallel_inf(self, integrand, res, spp, seed, rmin=(), rmax=(), bins=None, record=False, rec_cap=None):
        res = np.ascontiguousarray(np.asarray(res, dtype=np.uint64))
        nb = int(np.prod(res))
        rmin, rmax = _f32(rmin), _f32(rmax)
        bins = np.zeros(nb, np.float32) if bins is None else np.ascontiguousarray(bins, dtype=np.float32).copy()
        s1 = s2 = lens = elems = None
        used = ctypes.c_uint64(0)
        cap = 0
        if record == "moments":      # per-bin sum f, sum f^2 only (the one kind of record the multi-threaded build can take)
            s1 = np.zeros(nb, np.float64); s2 = np.zeros(nb, np.float64)
        elif record:
            s1 = np.zeros(nb, np.float64); s2 = np.zeros(nb, np.float64)
            lens = np.zeros(nb * spp, np.uint32)
            cap = rec_cap if rec_cap is not None else nb * spp * 64
            elems = np.zeros(cap, np.float32)
        rc = self.lib.vo_mc_per_bin_parallel_inf(integrand.encode(), len(res), _p(res), _p(rmin), _p(rmax), len(rmin),
                                                 ctypes.c_uint64(spp), ctypes.c_uint64(seed), _p(bins), _p(s1), _p(s2),
                                                 _p(lens), _p(elems), ctypes.c_uint64(cap), ctypes.byref(used))
        self._check(rc, "vo_mc_per_bin_parallel_inf")
        if record == "moments":
            return bins, s1, s2
        if record:
            return bins, s1, s2, lens, elems[: used.value]
        return bins

    def per_bin_parallel_mc_inf(self, integrand, res, spp, seed, rmin=(), rmax=(), bins=None, record=False, rec_cap=None):
        res = np.ascontiguousarray(np.asarray(res, dtype=np.uint64))
        nb = int(np.prod(res))
        rmin, rmax = _f32(rmin), _f32(rmax)
        bins = np.zeros(nb, np.float32) if bins is None else np.ascontiguousarray(bins, dtype=np.float32).copy()
        s1 = s2 = lens = elems = None
        used = ctypes.c_uint64(0)
        cap = 0
        if record:
            s1 = np.zeros(nb, np.float64); s2 = np.zeros(nb, np.float64)
            lens = np.zeros(nb * spp, np.uint32)
            cap = rec_cap if rec_cap is not None else nb * spp * 64
            elems = np.zeros(cap, np.float32)
        rc = self.lib.vo_per_bin_parallel_mc_inf(integrand.encode(), len(res), _p(res), _p(rmin), _p(rmax), len(rmin),
                                                 ctypes.c_uint64(spp), ctypes.c_uint64(seed), _p(bins), _p(s1), _p(s2),
                                                 _p(lens), _p(elems), ctypes.c_uint64(cap), ctypes.byref(used))
        self._check(rc, "vo_per_bin_parallel_mc_inf")
        if record:
            return bins, s1, s2, lens, elems[: used.value]
        return bins

    def monte_carlo_inf(self, integrand, res, samples, seed, rmin=(), rmax=(), bins=None):
        res = np.ascontiguousarray(np.asarray(res, dtype=np.uint64))
        rmin, rmax = _f32(rmin), _f32(rmax)
        bins = np.zeros(int(np.prod(res)), np.float32) if bins is None else np.ascontiguousarray(bins, dtype=np.float32).copy()
        rc = self.lib.vo_monte_carlo_inf(integrand.encode(), len(res), _p(res), _p(rmin), _p(rmax), len(rmin),
                                         ctypes.c_uint64(samples), ctypes.c_uint64(seed), _p(bins))
        self._check(rc, "vo_monte_carlo_inf")
        return bins

    def newton_cotes(self, integrand, rule, res, rmin, rmax, bins=None):
        res, rmin, rmax, nb = self._setup(res, rmin, rmax)
        bins = np.zeros(nb, np.float32) if bins is None else np.ascontiguousarray(bins, dtype=np.float32).copy()
        rc = self.lib.vo_newton_cotes(integrand.encode(), rule.encode(), len(res), _p(res), _p(rmin), _p(rmax), _p(bins))
        self._check(rc, "vo_newton_cotes")
        return bins

    def _region_buffers(self, integrand, rule, iterations):
        d = self.dim(integrand); n = iterations + 1; sd = RULE_SAMPLES[rule] ** d
        return dict(min=np.zeros((n, d), np.float32), max=np.zeros((n, d), np.float32), err=np.zeros(n, np.float32),
                    dim=np.zeros(n, np.uint32), data=np.zeros((n, sd), np.float32))

    def adaptive_iterations(self, integrand, rule, heuristic, iterations, res, rmin, rmax, size_weight=1e-5, bins=None):
        """returns (bins, regions) with regions = dict(min,max,err,dim,data) in the reference's heap-array order"""
        res, rmin, rmax, nb = self._setup(res, rmin, rmax)
        bins = np.zeros(nb, np.float32) if bins is None else np.ascontiguousarray(bins, dtype=np.float32).copy()
        reg = self._region_buffers(integrand, rule, iterations)
        rc = self.lib.vo_adaptive_iterations(integrand.encode(), rule.encode(), heuristic.encode(), ctypes.c_double(size_weight),
                                             ctypes.c_uint64(iterations), len(res), _p(res), _p(rmin), _p(rmax), _p(bins),
                                             _p(reg["min"]), _p(reg["max"]), _p(reg["err"]), _p(reg["dim"]), _p(reg["data"]))
        self._check(rc, "vo_adaptive_iterations")
        return bins, reg

    def adaptive_tolerance(self, integrand, rule, heuristic, tolerance, res, rmin, rmax, size_weight=1e-5, bins=None, reg_cap=0):
        """returns (bins, nleaves, regions or None); regions (leaves in the reference's depth-first visiting order) only from the port"""
        res, rmin, rmax, nb = self._setup(res, rmin, rmax)
        bins = np.zeros(nb, np.float32) if bins is None else np.ascontiguousarray(bins, dtype=np.float32).copy()
        d = self.dim(integrand); sd = RULE_SAMPLES[rule] ** d
        reg = None
        if reg_cap and self.kind == "port":
            reg = dict(min=np.zeros((reg_cap, d), np.float32), max=np.zeros((reg_cap, d), np.float32), err=np.zeros(reg_cap, np.float32),
                       dim=np.zeros(reg_cap, np.uint32), data=np.zeros((reg_cap, sd), np.float32))
        n = ctypes.c_uint64(0)
        self.lib.vo_adaptive_tolerance.restype = ctypes.c_int
        rc = self.lib.vo_adaptive_tolerance(integrand.encode(), rule.encode(), heuristic.encode(), ctypes.c_double(size_weight), ctypes.c_float(tolerance),
                                            len(res), _p(res), _p(rmin), _p(rmax), _p(bins), ctypes.byref(n), ctypes.c_uint64(reg_cap if reg else 0),
                                            _p(reg["min"]) if reg else None, _p(reg["max"]) if reg else None, _p(reg["err"]) if reg else None,
                                            _p(reg["dim"]) if reg else None, _p(reg["data"]) if reg else None)
        self._check(rc, "vo_adaptive_tolerance")
        if reg:
            reg = {k: v[: n.value] for k, v in reg.items()}
        return bins, int(n.value), reg

    # ---- double precision twins ----
    def newton_cotes_f64(self, integrand, rule, res, rmin, rmax, bins=None):
        res = np.ascontiguousarray(np.asarray(res, dtype=np.uint64)); nb = int(np.prod(res))
        rmin = np.ascontiguousarray(rmin, np.float64); rmax = np.ascontiguousarray(rmax, np.float64)
        bins = np.zeros(nb, np.float64) if bins is None else np.ascontiguousarray(bins, dtype=np.float64).copy()
        self.lib.vo_newton_cotes_f64.restype = ctypes.c_int
        rc = self.lib.vo_newton_cotes_f64(integrand.encode(), rule.encode(), len(res), _p(res), _p(rmin), _p(rmax), _p(bins))
        self._check(rc, "vo_newton_cotes_f64")
        return bins

    def adaptive_iterations_f64(self, integrand, rule, heuristic, iterations, res, rmin, rmax, size_weight=1e-5, bins=None):
        res = np.ascontiguousarray(np.asarray(res, dtype=np.uint64)); nb = int(np.prod(res))
        rmin = np.ascontiguousarray(rmin, np.float64); rmax = np.ascontiguousarray(rmax, np.float64)
        d = len(rmin); n = iterations + 1; sd = RULE_SAMPLES[rule] ** d
        bins = np.zeros(nb, np.float64) if bins is None else np.ascontiguousarray(bins, dtype=np.float64).copy()
        reg = dict(min=np.zeros((n, d), np.float64), max=np.zeros((n, d), np.float64), err=np.zeros(n, np.float64),
                   dim=np.zeros(n, np.uint32), data=np.zeros((n, sd), np.float64))
        self.lib.vo_adaptive_iterations_f64.restype = ctypes.c_int
        rc = self.lib.vo_adaptive_iterations_f64(integrand.encode(), rule.encode(), heuristic.encode(), ctypes.c_double(size_weight),
                                                 ctypes.c_uint64(iterations), len(res), _p(res), _p(rmin), _p(rmax), _p(bins),
                                                 _p(reg["min"]), _p(reg["max"]), _p(reg["err"]), _p(reg["dim"]), _p(reg["data"]))
        self._check(rc, "vo_adaptive_iterations_f64")
        return bins, reg

    def crespo2021(self, integrand, iterations, spp, seed, res, rmin, rmax, record=False):
        res, rmin, rmax, nb = self._setup(res, rmin, rmax)
        d = self.dim(integrand)
        bins = np.zeros(nb, np.float32)
        reg = self._region_buffers(integrand, "simpson_trapezoidal", iterations)
        nreg = approx = chosen = samples = None
        if record:
            nreg = np.zeros(nb, np.uint32); approx = np.zeros(nb, np.float32)
            chosen = np.zeros((nb, spp), np.uint32); samples = np.zeros((nb, spp, d), np.float32)
        rc = self.lib.vo_crespo2021(integrand.encode(), ctypes.c_uint64(iterations), ctypes.c_uint64(spp), ctypes.c_uint64(seed),
                                    len(res), _p(res), _p(rmin), _p(rmax), _p(bins), _p(nreg), _p(approx), _p(chosen), _p(samples),
                                    _p(reg["min"]), _p(reg["max"]), _p(reg["err"]), _p(reg["dim"]), _p(reg["data"]))
        self._check(rc, "vo_crespo2021")
        if record:
            return bins, reg, dict(nregions=nreg, approx=approx, chosen=chosen, samples=samples)
        return bins, reg

    # ---- Fubini family: first `nfirst` dims by the named first integrator, the rest by monte_carlo(mc_samples, mc_seed) ----
    def _fubini_args(self, integrand, res, rmin, rmax):
        res = np.ascontiguousarray(np.asarray(res, dtype=np.uint64))
        d = self.dim(integrand)
        if d > 0 and len(rmin) == 0:
            rmin, rmax = [0.0] * d, [1.0] * d
        rmin, rmax = _f32(rmin), _f32(rmax)
        return res, rmin, rmax, np.zeros(int(np.prod(res)), np.float32)

    def fubini_adaptive_mc(self, integrand, nfirst, rule, heuristic, iterations, mc_samples, mc_seed, res, rmin=(), rmax=(), size_weight=1e-5):
        res, rmin, rmax, bins = self._fubini_args(integrand, res, rmin, rmax)
        self.lib.vo_fubini_adaptive_mc.restype = ctypes.c_int
        rc = self.lib.vo_fubini_adaptive_mc(integrand.encode(), int(nfirst), rule.encode(), heuristic.encode(), ctypes.c_double(size_weight),
                                            ctypes.c_uint64(iterations), ctypes.c_uint64(mc_samples), ctypes.c_uint64(mc_seed),
                                            len(res), _p(res), _p(rmin), _p(rmax), len(rmin), _p(bins))
        self._check(rc, "vo_fubini_adaptive_mc")
        return bins

    def fubini_mc_mc(self, integrand, nfirst, spp, seed, mc_samples, mc_seed, res, rmin=(), rmax=()):
        res, rmin, rmax, bins = self._fubini_args(integrand, res, rmin, rmax)
        self.lib.vo_fubini_mc_mc.restype = ctypes.c_int
        rc = self.lib.vo_fubini_mc_mc(integrand.encode(), int(nfirst), ctypes.c_uint64(spp), ctypes.c_uint64(seed), ctypes.c_uint64(mc_samples),
                                      ctypes.c_uint64(mc_seed), len(res), _p(res), _p(rmin), _p(rmax), len(rmin), _p(bins))
        self._check(rc, "vo_fubini_mc_mc")
        return bins

    def crespo2021_infinite(self, integrand, nfirst, iterations, mc_samples, spp, seed, res, rmin=(), rmax=()):
        res, rmin, rmax, bins = self._fubini_args(integrand, res, rmin, rmax)
        self.lib.vo_crespo2021_infinite.restype = ctypes.c_int
        rc = self.lib.vo_crespo2021_infinite(integrand.encode(), int(nfirst), ctypes.c_uint64(iterations), ctypes.c_uint64(mc_samples),
                                             ctypes.c_uint64(spp), ctypes.c_uint64(seed), len(res), _p(res), _p(rmin), _p(rmax), len(rmin), _p(bins))
        self._check(rc, "vo_crespo2021_infinite")
        return bins

    def cv_fixed_weight(self, integrand, iterations, spp, seed, alpha, res, rmin, rmax, record=False):
        res, rmin, rmax, nb = self._setup(res, rmin, rmax)
        d = self.dim(integrand)
        bins = np.zeros(nb, np.float32)
        nreg = approx = chosen = samples = None
        if record:
            nreg = np.zeros(nb, np.uint32); approx = np.zeros(nb, np.float32)
            chosen = np.zeros((nb, spp), np.uint32); samples = np.zeros((nb, spp, d), np.float32)
        self.lib.vo_cv_fixed_weight.restype = ctypes.c_int
        rc = self.lib.vo_cv_fixed_weight(integrand.encode(), ctypes.c_uint64(iterations), ctypes.c_uint64(spp), ctypes.c_uint64(seed), ctypes.c_double(alpha),
                                         len(res), _p(res), _p(rmin), _p(rmax), _p(bins), _p(nreg), _p(approx), _p(chosen), _p(samples))
        self._check(rc, "vo_cv_fixed_weight")
        if record:
            return bins, dict(nregions=nreg, approx=approx, chosen=chosen, samples=samples)
        return bins

    RR_POLICIES = {"uniform": 0, "integral": 1, "error": 2, "pdf": 3}

    def cv_policies(self, integrand, iterations, spp, seed, rr, res, rmin, rmax, fixed_alpha=None, record=False):
        """integrator_adaptive_variance_reduction_parallel(nested(simpson,trapezoidal), size/relative 1e-5, iterations, rr_<rr>_region(),
        cv_optimize_weight() | cv_fixed_weight(fixed_alpha), region_sampling_uniform(), spp, seed)"""
        res, rmin, rmax, nb = self._setup(res, rmin, rmax)
        d = self.dim(integrand)
        bins = np.zeros(nb, np.float32)
        nreg = approx = chosen = samples = None
        if record:
            nreg = np.zeros(nb, np.uint32); approx = np.zeros(nb, np.float32)
            chosen = np.zeros((nb, spp), np.uint32); samples = np.zeros((nb, spp, d), np.float32)
        self.lib.vo_cv_policies.restype = ctypes.c_int
        rc = self.lib.vo_cv_policies(integrand.encode(), ctypes.c_uint64(iterations), ctypes.c_uint64(spp), ctypes.c_uint64(seed),
                                     ctypes.c_int(self.RR_POLICIES[rr]), ctypes.c_int(0 if fixed_alpha is None else 1), ctypes.c_double(1.0 if fixed_alpha is None else fixed_alpha),
                                     len(res), _p(res), _p(rmin), _p(rmax), _p(bins), _p(nreg), _p(approx), _p(chosen), _p(samples))
        self._check(rc, "vo_cv_policies")
        if record:
            return bins, dict(nregions=nreg, approx=approx, chosen=chosen, samples=samples)
        return bins

    def cv_sampling(self, integrand, iterations, spp, seed, rs, res, rmin, rmax, power=1.0, cutoff=0.0):
        """crespo2021 pipeline with region_sampling_<rs> (reference builds only)"""
        res, rmin, rmax, nb = self._setup(res, rmin, rmax)
        bins = np.zeros(nb, np.float32)
        self.lib.vo_cv_sampling.restype = ctypes.c_int
        rc = self.lib.vo_cv_sampling(integrand.encode(), ctypes.c_uint64(iterations), ctypes.c_uint64(spp), ctypes.c_uint64(seed),
                                     ctypes.c_int({"uniform": 0, "importance": 1, "mis": 2, "russian_roulette": 3}[rs]), ctypes.c_double(power), ctypes.c_double(cutoff),
                                     len(res), _p(res), _p(rmin), _p(rmax), _p(bins))
        self._check(rc, "vo_cv_sampling")
        return bins

    def cv_optimized_infinite(self, integrand, nfirst, iterations, mc_samples, spp, seed, res, rmin=(), rmax=()):
        """integrator_adaptive_fubini_variance_reduction_parallel_optimized<nfirst> (reference builds only)"""
        res, rmin, rmax, bins = self._fubini_args(integrand, res, rmin, rmax)
        self.lib.vo_cv_optimized_infinite.restype = ctypes.c_int
        rc = self.lib.vo_cv_optimized_infinite(integrand.encode(), int(nfirst), ctypes.c_uint64(iterations), ctypes.c_uint64(mc_samples),
                                               ctypes.c_uint64(spp), ctypes.c_uint64(seed), len(res), _p(res), _p(rmin), _p(rmax), len(rmin), _p(bins))
        self._check(rc, "vo_cv_optimized_infinite")
        return bins

    def mt_per_bin(self, path, integrand, res, spp, seed, nthreads, rmin=(), rmax=()):
        res = np.ascontiguousarray(np.asarray(res, dtype=np.uint64))
        rmin, rmax = _f32(rmin), _f32(rmax)
        bins = np.zeros(int(np.prod(res)), np.float32)
        rc = self.lib.vo_mt_per_bin(path.encode(), integrand.encode(), len(res), _p(res), _p(rmin), _p(rmax), len(rmin),
                                    ctypes.c_uint64(spp), ctypes.c_uint64(seed), int(nthreads), _p(bins))
        self._check(rc, "vo_mt_per_bin")
        return bins


_cache = {}


def load(kind="port"):
    """kind: 'port' (restatement, built on demand) or 'reference' (prebuilt .so, or built when /root/reference exists)."""
    if kind not in _cache:
        path = _PATHS[kind]
        if not os.path.exists(path):
            build(kind)
        _cache[kind] = Oracle(path)
    return _cache[kind]
