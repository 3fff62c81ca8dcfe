"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.json")


def load_golden():
    with open(GOLDEN) as f:
        return json.load(f)["vectors"]


def f32(bits):
    return np.asarray(bits, dtype=np.uint32).view(np.float32)


def f64(bits):
    return np.asarray(bits, dtype=np.uint64).view(np.float64)


def same_bits(a, b):
    a = np.ascontiguousarray(a).ravel()
    b = np.ascontiguousarray(b).ravel()
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def assert_same_bits(a, b, what=""):
    a = np.ascontiguousarray(a).ravel()
    b = np.ascontiguousarray(b).ravel()
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if not same_bits(a, b):
        bad = np.flatnonzero(a != b)
        raise AssertionError(f"{what}: {bad.size}/{a.size} elements differ, first at {bad[:5]}: {a[bad[:5]]} vs {b[bad[:5]]}")


def data_checksum(data):
    return int(np.sum(np.ascontiguousarray(data).view(np.uint32).astype(np.uint64)) & 0xFFFFFFFFFFFF)


def run_oracle_vector(O, v):
    """Run vector `v` (a dict from reference_vectors.json) through oracle `O`; return dict of outputs keyed like the vector."""
    integ, res, rmin, rmax, path = v["integrand"], v["res"], v["rmin"], v["rmax"], v["path"]
    if path in ("mc_per_bin_parallel", "per_bin_parallel_mc"):
        b, s, s1, s2 = getattr(O, path)(integ, res, rmin, rmax, v["spp"], v["seed"], record=True)
        return dict(bins=b, samples=s, sum=s1, sum2=s2)
    if path == "monte_carlo":
        b, s = O.monte_carlo(integ, res, rmin, rmax, v["samples_n"], v["seed"], record=True)
        return dict(bins=b, samples=s)
    if path == "newton_cotes":
        return dict(bins=O.newton_cotes(integ, v["rule"], res, rmin, rmax))
    if path == "adaptive_iterations":
        b, reg = O.adaptive_iterations(integ, v["rule"], v["heuristic"], v["iterations"], res, rmin, rmax, v["size_weight"])
        return dict(bins=b, reg_min=reg["min"], reg_max=reg["max"], reg_err=reg["err"], reg_dim=reg["dim"],
                    reg_data_checksum=data_checksum(reg["data"]))
    if path == "crespo2021":
        b, reg, rec = O.crespo2021(integ, v["iterations"], v["spp"], v["seed"], res, rmin, rmax, record=True)
        return dict(bins=b, nregions=rec["nregions"], approx=rec["approx"], chosen=rec["chosen"], samples=rec["samples"])
    if path == "mc_per_bin_parallel_inf":
        b, s1, s2, lens, elems = O.mc_per_bin_parallel_inf(integ, res, v["spp"], v["seed"], rmin, rmax, record=True)
        return dict(bins=b, sum=s1, sum2=s2, lens=lens, elems=elems)
    if path == "per_bin_parallel_mc_inf":
        b, s1, s2, lens, elems = O.per_bin_parallel_mc_inf(integ, res, v["spp"], v["seed"], rmin, rmax, record=True)
        return dict(bins=b, sum=s1, sum2=s2, lens=lens, elems=elems)
    if path == "monte_carlo_inf":
        return dict(bins=O.monte_carlo_inf(integ, res, v["samples_n"], v["seed"], rmin, rmax))
    if path == "cv_fixed_weight":
        b, rec = O.cv_fixed_weight(integ, v["iterations"], v["spp"], v["seed"], v["alpha"], res, rmin, rmax, record=True)
        return dict(bins=b, nregions=rec["nregions"], chosen=rec["chosen"], samples=rec["samples"])
    if path == "cv_policies":
        b, rec = O.cv_policies(integ, v["iterations"], v["spp"], v["seed"], v["rr"], res, rmin, rmax, fixed_alpha=v["alpha"], record=True)
        return dict(bins=b, nregions=rec["nregions"], chosen=rec["chosen"], samples=rec["samples"])
    if path == "adaptive_tolerance":
        b, n, _ = O.adaptive_tolerance(integ, v["rule"], v["heuristic"], v["tolerance"], res, rmin, rmax, v["size_weight"])
        return dict(bins=b, nleaves=n)
    if path == "fubini_adaptive_mc":
        return dict(bins=O.fubini_adaptive_mc(integ, v["nfirst"], v["rule"], v["heuristic"], v["iterations"], v["mc_samples"], v["mc_seed"], res, rmin, rmax, v["size_weight"]))
    if path == "fubini_mc_mc":
        return dict(bins=O.fubini_mc_mc(integ, v["nfirst"], v["spp"], v["seed"], v["mc_samples"], v["mc_seed"], res, rmin, rmax))
    if path == "crespo2021_infinite":
        return dict(bins=O.crespo2021_infinite(integ, v["nfirst"], v["iterations"], v["mc_samples"], v["spp"], v["seed"], res, rmin, rmax))
    raise KeyError(path)


def golden_field(v, key):
    x = v[key]
    if key in ("sum", "sum2"):
        return f64(x)
    if key in ("reg_dim", "nregions", "chosen", "lens"):
        return np.asarray(x, dtype=np.uint32)
    if key in ("reg_data_checksum", "nleaves"):
        return x
    return f32(x)
