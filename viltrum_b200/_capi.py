"""ctypes binding of the C ABI in include/viltrum_b200.h (libviltrum_b200.so).  No torch types cross this
boundary: pointers are integers / numpy buffers, sizes are plain ints.  The library has no CPU path — importing
works anywhere, creating a Context needs a CUDA device and raises otherwise."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libviltrum_b200.so")

MAX_DIM, MAX_DIMBINS, K_COUNT = 8, 3, 8
HOST, DEVICE = 0, 1
MC_PER_BIN, PER_BIN_MC = 0, 1
MC_RNG_PHILOX, MC_LATTICE24, MC_ALLREDUCE = 1, 2, 4      # vb200_mc_params.options
COMM_ID_BYTES = 128
SHARD_EMPTY_INDEX = 0xFFFFFFFFFFFFFFFF   # VB200_SHARD_EMPTY_INDEX: {EMPTY, EMPTY} = explicit empty shard ({0,0} = whole grid)


def shard_pair(shard):
    """(begin, end) for a vb200_shard: None -> {0,0} (whole grid); an empty range -> the explicit empty encoding"""
    if not shard:
        return 0, 0
    b, e = int(shard[0]), int(shard[1])
    return (SHARD_EMPTY_INDEX, SHARD_EMPTY_INDEX) if b == e else (b, e)
CV_OPTIMIZE_WEIGHT, CV_FIXED_WEIGHT = 0, 1
RR_POLICIES = {"uniform": 0, "integral": 1, "error": 2, "pdf": 3, "stratified": 4}     # vb200_rr_policy
RS_POLICIES = {"uniform": 0, "importance": 1, "mis": 2, "russian_roulette": 3}          # vb200_rs_policy
RULES = {"trapezoidal": 2, "simpson": 3, "boole": 5, "simpson_trapezoidal": 32, "boole_simpson": 53}
RULE_SAMPLES = {2: 2, 3: 3, 5: 5, 32: 3, 53: 5}


def rule_id(name):
    """'simpson', 'boole_simpson', ... or 'steps<N>_<rule>' (Steps<Q,N>, reference src/newton-cotes/rules.h:321-388) -> vb200_rule code"""
    if isinstance(name, int):
        return name
    if name.startswith("steps"):
        n, base = name[5:].split("_", 1)
        return 0x1000000 | (RULES[base] << 16) | (int(n) & 0xffff)
    return RULES[name]
HEURISTICS = {"default": 0, "size": 1, "mixed": 2}
METRICS = {"absolute": 0, "relative": 1}

STATUS = {0: "VB200_OK", -1: "VB200_ERR_NO_DEVICE", -2: "VB200_ERR_INVALID", -3: "VB200_ERR_CUDA",
          -4: "VB200_ERR_UNSUPPORTED", -5: "VB200_ERR_NOMEM"}

# every symbol include/viltrum_b200.h declares (tests/test_capi_symbols.py checks the header against this list and the .so)
SYMBOLS = [
    "vb200_create", "vb200_destroy", "vb200_last_error", "vb200_stream", "vb200_synchronize", "vb200_sm_count",
    "vb200_launch_count", "vb200_kernel_timer", "vb200_kernel_timer_read", "vb200_host_register", "vb200_host_unregister", "vb200_measure_fp32_peak", "vb200_philox4x32_10", "vb200_xoshiro128pp", "vb200_threefry4x32", "vb200_builtin_integrand", "vb200_builtin_count", "vb200_builtin_name",
    "vb200_mc_per_bin", "vb200_mc_per_bin_replay", "vb200_mc_per_bin_inf", "vb200_mc_per_bin_inf_replay",
    "vb200_monte_carlo", "vb200_regions_generate_adaptive", "vb200_regions_generate_single", "vb200_regions_upload",
    "vb200_regions_count", "vb200_regions_dim", "vb200_regions_samples", "vb200_regions_download", "vb200_regions_free",
    "vb200_regions_integrate_bins", "vb200_cv_integrate", "vb200_cv_replay",
    "vb200_regions_generate_single_f64", "vb200_regions_upload_f64", "vb200_regions_download_f64", "vb200_regions_integrate_bins_f64",
    "vb200_builtin_integrand_f64", "vb200_builtin_fubini", "vb200_integrand_free", "vb200_regions_generate_tolerance",
    "vb200_regions_generate_adaptive_f64",
    "vb200_comm_unique_id", "vb200_comm_init", "vb200_comm_destroy", "vb200_comm_rank", "vb200_comm_size", "vb200_nccl_version", "vb200_regions_broadcast",
]


class Domain(ctypes.Structure):
    _fields_ = [("dim", ctypes.c_int32), ("dimbins", ctypes.c_int32), ("rmin", ctypes.c_float * MAX_DIM),
                ("rmax", ctypes.c_float * MAX_DIM), ("res", ctypes.c_uint64 * MAX_DIMBINS),
                ("drange", ctypes.c_float * MAX_DIMBINS), ("reserved", ctypes.c_int32)]


class Domain64(ctypes.Structure):
    _fields_ = [("dim", ctypes.c_int32), ("dimbins", ctypes.c_int32), ("rmin", ctypes.c_double * MAX_DIM),
                ("rmax", ctypes.c_double * MAX_DIM), ("res", ctypes.c_uint64 * MAX_DIMBINS)]


class Shard(ctypes.Structure):
    _fields_ = [("begin", ctypes.c_uint64), ("end", ctypes.c_uint64)]


class McParams(ctypes.Structure):
    _fields_ = [("domain", Domain), ("shard", Shard), ("spp", ctypes.c_uint64), ("seed", ctypes.c_uint64),
                ("flavor", ctypes.c_int32), ("options", ctypes.c_int32)]


class MixedHeuristic(ctypes.Structure):
    _fields_ = [("metric_rest", ctypes.c_int32), ("dimension", ctypes.c_int32), ("bins_weight", ctypes.c_double), ("size_threshold_bins", ctypes.c_double),
                ("size_threshold_rest", ctypes.c_double), ("error_increase_factor", ctypes.c_double)]


class AdaptiveParams(ctypes.Structure):
    _fields_ = [("domain", Domain), ("rule", ctypes.c_int32), ("heuristic", ctypes.c_int32), ("metric", ctypes.c_int32),
                ("batch", ctypes.c_int32), ("size_weight", ctypes.c_double), ("iterations", ctypes.c_uint64), ("mixed", MixedHeuristic)]


class AdaptiveParams64(ctypes.Structure):
    _fields_ = [("domain", Domain64), ("rule", ctypes.c_int32), ("heuristic", ctypes.c_int32), ("metric", ctypes.c_int32),
                ("batch", ctypes.c_int32), ("size_weight", ctypes.c_double), ("iterations", ctypes.c_uint64), ("mixed", MixedHeuristic)]


class ToleranceParams(ctypes.Structure):
    _fields_ = [("domain", Domain), ("rule", ctypes.c_int32), ("heuristic", ctypes.c_int32), ("metric", ctypes.c_int32),
                ("tolerance", ctypes.c_float), ("size_weight", ctypes.c_double), ("max_regions", ctypes.c_uint64)]


class CvParams(ctypes.Structure):
    _fields_ = [("domain", Domain), ("shard", Shard), ("spp", ctypes.c_uint64), ("seed", ctypes.c_uint64),
                ("weight_strategy", ctypes.c_int32), ("rr_policy", ctypes.c_int32), ("alpha", ctypes.c_double),
                ("rs_policy", ctypes.c_int32), ("reserved", ctypes.c_int32), ("rs_power", ctypes.c_double), ("rs_cutoff", ctypes.c_double)]


class Vb200Error(RuntimeError):
    def __init__(self, status, text):
        super().__init__(f"{STATUS.get(status, status)}: {text}")
        self.status = status


_lib = None


def lib():
    """Loads libviltrum_b200.so (built in-tree by viltrum_b200.build).  Fails loudly if it is missing: there is no
    fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -m viltrum_b200.build` (no fallback path exists)")
        L = ctypes.CDLL(LIB_PATH)
        vp, i32, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64
        L.vb200_create.argtypes = [i32, ctypes.POINTER(vp)]; L.vb200_create.restype = i32
        L.vb200_destroy.argtypes = [vp]; L.vb200_destroy.restype = None
        L.vb200_last_error.argtypes = [vp]; L.vb200_last_error.restype = ctypes.c_char_p
        L.vb200_stream.argtypes = [vp]; L.vb200_stream.restype = vp
        L.vb200_synchronize.argtypes = [vp]; L.vb200_synchronize.restype = i32
        L.vb200_sm_count.argtypes = [vp]; L.vb200_sm_count.restype = i32
        L.vb200_launch_count.argtypes = [vp]; L.vb200_launch_count.restype = u64
        L.vb200_kernel_timer.argtypes = [vp, i32]; L.vb200_kernel_timer.restype = i32
        L.vb200_kernel_timer_read.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)]; L.vb200_kernel_timer_read.restype = i32
        L.vb200_host_register.argtypes = [vp, vp, ctypes.c_size_t]; L.vb200_host_register.restype = i32
        L.vb200_host_unregister.argtypes = [vp, vp]; L.vb200_host_unregister.restype = i32
        L.vb200_measure_fp32_peak.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_double)]; L.vb200_measure_fp32_peak.restype = i32
        L.vb200_philox4x32_10.argtypes = [ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]; L.vb200_philox4x32_10.restype = None
        L.vb200_xoshiro128pp.argtypes = [ctypes.POINTER(ctypes.c_uint32), u64, ctypes.POINTER(ctypes.c_uint32)]; L.vb200_xoshiro128pp.restype = None
        L.vb200_threefry4x32.argtypes = [i32, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]; L.vb200_threefry4x32.restype = i32
        L.vb200_comm_unique_id.argtypes = [vp, vp]; L.vb200_comm_unique_id.restype = i32
        L.vb200_comm_init.argtypes = [vp, vp, i32, i32]; L.vb200_comm_init.restype = i32
        L.vb200_comm_destroy.argtypes = [vp]; L.vb200_comm_destroy.restype = i32
        L.vb200_comm_rank.argtypes = [vp]; L.vb200_comm_rank.restype = i32
        L.vb200_comm_size.argtypes = [vp]; L.vb200_comm_size.restype = i32
        L.vb200_nccl_version.argtypes = []; L.vb200_nccl_version.restype = i32
        L.vb200_regions_broadcast.argtypes = [vp, ctypes.POINTER(vp), i32]; L.vb200_regions_broadcast.restype = i32
        L.vb200_builtin_integrand.argtypes = [ctypes.c_char_p, i32]; L.vb200_builtin_integrand.restype = vp
        L.vb200_builtin_count.argtypes = []; L.vb200_builtin_count.restype = i32
        L.vb200_builtin_name.argtypes = [i32]; L.vb200_builtin_name.restype = ctypes.c_char_p
        L.vb200_mc_per_bin.argtypes = [vp, vp, ctypes.POINTER(McParams), vp, i32, vp, vp]; L.vb200_mc_per_bin.restype = i32
        L.vb200_mc_per_bin_replay.argtypes = [vp, vp, ctypes.POINTER(McParams), vp, i32, vp, i32]; L.vb200_mc_per_bin_replay.restype = i32
        L.vb200_mc_per_bin_inf.argtypes = [vp, vp, ctypes.POINTER(McParams), vp, i32, vp, vp]; L.vb200_mc_per_bin_inf.restype = i32
        L.vb200_mc_per_bin_inf_replay.argtypes = [vp, vp, ctypes.POINTER(McParams), vp, vp, i32, vp, i32]; L.vb200_mc_per_bin_inf_replay.restype = i32
        L.vb200_monte_carlo.argtypes = [vp, vp, ctypes.POINTER(McParams), vp, i32]; L.vb200_monte_carlo.restype = i32
        L.vb200_regions_generate_adaptive.argtypes = [vp, vp, ctypes.POINTER(AdaptiveParams), ctypes.POINTER(vp)]; L.vb200_regions_generate_adaptive.restype = i32
        L.vb200_regions_generate_adaptive_f64.argtypes = [vp, vp, ctypes.POINTER(AdaptiveParams64), ctypes.POINTER(vp)]; L.vb200_regions_generate_adaptive_f64.restype = i32
        L.vb200_regions_generate_tolerance.argtypes = [vp, vp, ctypes.POINTER(ToleranceParams), ctypes.POINTER(vp)]; L.vb200_regions_generate_tolerance.restype = i32
        L.vb200_regions_generate_single.argtypes = [vp, vp, ctypes.POINTER(Domain), i32, ctypes.POINTER(vp)]; L.vb200_regions_generate_single.restype = i32
        L.vb200_regions_upload.argtypes = [vp, i32, i32, u64, vp, vp, vp, vp, vp, ctypes.POINTER(vp)]; L.vb200_regions_upload.restype = i32
        L.vb200_regions_count.argtypes = [vp]; L.vb200_regions_count.restype = u64
        L.vb200_regions_dim.argtypes = [vp]; L.vb200_regions_dim.restype = i32
        L.vb200_regions_samples.argtypes = [vp]; L.vb200_regions_samples.restype = i32
        L.vb200_regions_download.argtypes = [vp, vp, vp, vp, vp, vp, vp]; L.vb200_regions_download.restype = i32
        L.vb200_regions_free.argtypes = [vp]; L.vb200_regions_free.restype = None
        L.vb200_regions_integrate_bins.argtypes = [vp, vp, ctypes.POINTER(Domain), ctypes.POINTER(Shard), vp, i32]; L.vb200_regions_integrate_bins.restype = i32
        L.vb200_regions_generate_single_f64.argtypes = [vp, vp, ctypes.POINTER(Domain64), i32, ctypes.POINTER(vp)]; L.vb200_regions_generate_single_f64.restype = i32
        L.vb200_regions_upload_f64.argtypes = [vp, i32, i32, u64, vp, vp, vp, vp, vp, ctypes.POINTER(vp)]; L.vb200_regions_upload_f64.restype = i32
        L.vb200_regions_download_f64.argtypes = [vp, vp, vp, vp, vp, vp, vp]; L.vb200_regions_download_f64.restype = i32
        L.vb200_regions_integrate_bins_f64.argtypes = [vp, vp, ctypes.POINTER(Domain64), ctypes.POINTER(Shard), vp, i32]; L.vb200_regions_integrate_bins_f64.restype = i32
        L.vb200_builtin_integrand_f64.argtypes = [ctypes.c_char_p, i32]; L.vb200_builtin_integrand_f64.restype = vp
        L.vb200_builtin_fubini.argtypes = [ctypes.c_char_p, i32, vp, vp, i32, u64, u64]; L.vb200_builtin_fubini.restype = vp
        L.vb200_integrand_free.argtypes = [vp]; L.vb200_integrand_free.restype = None
        L.vb200_cv_integrate.argtypes = [vp, vp, vp, ctypes.POINTER(CvParams), vp, i32, vp, vp]; L.vb200_cv_integrate.restype = i32
        L.vb200_cv_replay.argtypes = [vp, vp, vp, ctypes.POINTER(CvParams), vp, vp, i32, vp, i32]; L.vb200_cv_replay.restype = i32
        _lib = L
    return _lib


def make_domain64(dim, res, rmin, rmax):
    d = Domain64()
    d.dim = int(dim); d.dimbins = len(res)
    for i, r in enumerate(res):
        d.res[i] = int(r)
    for i in range(MAX_DIM):
        d.rmin[i] = float(rmin[i]) if i < len(rmin) else 0.0
        d.rmax[i] = float(rmax[i]) if i < len(rmax) else 1.0
    return d


def make_domain(dim, res, rmin=(), rmax=()):
    d = Domain()
    d.dim = int(dim); d.dimbins = len(res)
    if len(res) > MAX_DIMBINS:
        raise ValueError(f"at most {MAX_DIMBINS} binned dimensions")
    for i, r in enumerate(res):
        d.res[i] = int(r)
    for i in range(MAX_DIM):
        d.rmin[i] = float(rmin[i]) if i < len(rmin) else 0.0
        d.rmax[i] = float(rmax[i]) if i < len(rmax) else 1.0
    return d
