"""Generates tests/golden/reference_vectors.json by running the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run in the authoring container:  python tests/golden/generate_golden.py

Floats are stored as uint32 bit patterns so the comparison is bit-exact.  The reference has no golden vectors
of its own; these pin the oracle restatement and, through it, the CUDA exact modes on machines
where /root/reference does not exist.
"""
import json
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import pyoracle  # noqa: E402


def bits(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        return a.view(np.uint32).ravel().tolist()
    if a.dtype == np.float64:
        return a.view(np.uint64).ravel().tolist()
    return a.ravel().tolist()


CASES = [  # (integrand, res, rmin, rmax)
    ("x2y2", [5], 0.0, 1.0),
    ("x2y2", [4, 3], 0.0, 1.0),
    ("ind2", [6, 5], 0.0, 1.0),
    ("smooth_edge2", [8, 8], 0.0, 1.0),
    ("cubic1", [7], -0.5, 1.25),
    ("poly3", [3, 2], 0.1, 0.9),
    ("shade4_16", [4, 4], 0.0, 1.0),
    ("shade4_64", [3, 2], 0.0, 1.0),
    ("shade5_16", [3, 4], 0.0, 1.0),
    ("shade5_64", [2, 2], 0.0, 1.0),
]


def main():
    R = pyoracle.load("reference")
    assert R.kind == "reference"
    out = {"generator": "tests/golden/generate_golden.py", "source": "unmodified reference via oracle/_ref/libviltrum_ref.so",
           "flags": "g++ -std=c++17 -O3 -march=x86-64-v3 -ffp-contract=off", "encoding": "float32 as uint32 bits, float64 as uint64 bits",
           "vectors": []}
    V = out["vectors"]
    for integ, res, lo, hi in CASES:
        d = R.dim(integ)
        rmin, rmax = [lo] * d, [hi] * d
        base = dict(integrand=integ, res=res, rmin=rmin, rmax=rmax)
        for fn in ("mc_per_bin_parallel", "per_bin_parallel_mc"):
            b, s, s1, s2 = getattr(R, fn)(integ, res, rmin, rmax, 8, 3, record=True)
            V.append(dict(base, path=fn, spp=8, seed=3, bins=bits(b), samples=bits(s), sum=bits(s1), sum2=bits(s2)))
        b, s = R.monte_carlo(integ, res, rmin, rmax, 256, 5, record=True)
        V.append(dict(base, path="monte_carlo", samples_n=256, seed=5, bins=bits(b), samples=bits(s)))
        for rule in ("trapezoidal", "simpson", "boole"):
            if d >= 4 and rule == "boole":
                continue
            V.append(dict(base, path="newton_cotes", rule=rule, bins=bits(R.newton_cotes(integ, rule, res, rmin, rmax))))
        for rule in ("simpson_trapezoidal", "boole_simpson"):
            if d >= 4 and rule == "boole_simpson":
                continue
            for h in ("default_absolute", "size_relative"):
                it = 24 if d < 4 else 10
                b, reg = R.adaptive_iterations(integ, rule, h, it, res, rmin, rmax)
                V.append(dict(base, path="adaptive_iterations", rule=rule, heuristic=h, iterations=it, size_weight=1e-5, bins=bits(b),
                              reg_min=bits(reg["min"]), reg_max=bits(reg["max"]), reg_err=bits(reg["err"]), reg_dim=bits(reg["dim"]),
                              reg_data_checksum=int(np.sum(reg["data"].view(np.uint32).astype(np.uint64)) & 0xFFFFFFFFFFFF)))
        it = 12 if d < 4 else 6
        b, reg, rec = R.crespo2021(integ, it, 4, 9, res, rmin, rmax, record=True)
        V.append(dict(base, path="crespo2021", iterations=it, spp=4, seed=9, bins=bits(b), nregions=bits(rec["nregions"]),
                      approx=bits(rec["approx"]), chosen=bits(rec["chosen"]), samples=bits(rec["samples"])))
    for integ in ("walk", "decay"):
        for res, rmin, rmax in (([4], [], []), ([3, 2], [0.1, 0.2, 0.0], [0.9, 0.7, 1.0])):
            b, s1, s2, lens, elems = R.mc_per_bin_parallel_inf(integ, res, 6, 5, rmin, rmax, record=True)
            V.append(dict(integrand=integ, res=res, rmin=rmin, rmax=rmax, path="mc_per_bin_parallel_inf", spp=6, seed=5,
                          bins=bits(b), sum=bits(s1), sum2=bits(s2), lens=bits(lens), elems=bits(elems)))
            b, s1, s2, lens, elems = R.per_bin_parallel_mc_inf(integ, res, 6, 5, rmin, rmax, record=True)
            V.append(dict(integrand=integ, res=res, rmin=rmin, rmax=rmax, path="per_bin_parallel_mc_inf", spp=6, seed=5,
                          bins=bits(b), sum=bits(s1), sum2=bits(s2), lens=bits(lens), elems=bits(elems)))
            V.append(dict(integrand=integ, res=res, rmin=rmin, rmax=rmax, path="monte_carlo_inf", samples_n=300, seed=7,
                          bins=bits(R.monte_carlo_inf(integ, res, 300, 7, rmin, rmax))))
    # Steps<Q,N> composite rules: integrator_newton_cotes(steps<N>(rule))
    for integ, res, lo, hi in (("x2y2", [5], 0.0, 1.0), ("x2y2", [4, 3], 0.05, 1.1), ("smooth_edge2", [8, 8], 0.0, 1.0), ("cubic1", [7], -0.5, 1.25),
                               ("poly3", [3, 2], 0.1, 0.9), ("ind2", [6, 5], 0.0, 1.0), ("shade4_16", [3, 4], 0.0, 1.0)):
        d = R.dim(integ)
        for rule in ("steps2_boole", "steps3_simpson", "steps4_trapezoidal", "steps16_trapezoidal"):
            if d >= 4 and rule == "steps16_trapezoidal":
                continue
            V.append(dict(integrand=integ, res=res, rmin=[lo] * d, rmax=[hi] * d, path="newton_cotes", rule=rule, bins=bits(R.newton_cotes(integ, rule, res, [lo] * d, [hi] * d))))
    # cv_fixed_weight: bins + the recorded region choices / sample points for the replay mode
    for integ, res, it, spp, alpha in (("x2y2", [5], 12, 6, 1.0), ("smooth_edge2", [6, 6], 30, 4, 0.5), ("shade4_16", [3, 3], 8, 4, 0.0), ("poly3", [3, 2], 12, 4, 0.75)):
        d = R.dim(integ)
        b, rec = R.cv_fixed_weight(integ, it, spp, 9, alpha, res, [0.0] * d, [1.0] * d, record=True)
        V.append(dict(integrand=integ, res=res, rmin=[0.0] * d, rmax=[1.0] * d, path="cv_fixed_weight", iterations=it, spp=spp, seed=9, alpha=alpha,
                      bins=bits(b), nregions=bits(rec["nregions"]), chosen=bits(rec["chosen"]), samples=bits(rec["samples"])))
    # rr_integral_region / rr_error_region with either weight strategy: bins + recorded choices / sample points
    for integ, res, it, spp, rr, alpha in (("x2y2", [5], 12, 6, "integral", None), ("smooth_edge2", [6, 6], 30, 4, "error", None), ("shade4_16", [3, 3], 8, 4, "integral", 0.0),
                                           ("poly3", [3, 2], 12, 4, "error", 1.0), ("ind2", [5, 4], 20, 5, "integral", 0.5), ("shade5_16", [2, 3], 6, 3, "error", None),
                                           ("smooth_edge2", [5, 5], 25, 4, "pdf", None), ("shade4_16", [3, 2], 10, 5, "pdf", 0.0), ("cubic1", [6], 9, 4, "pdf", 1.0)):
        d = R.dim(integ)
        b, rec = R.cv_policies(integ, it, spp, 9, rr, res, [0.0] * d, [1.0] * d, fixed_alpha=alpha, record=True)
        V.append(dict(integrand=integ, res=res, rmin=[0.0] * d, rmax=[1.0] * d, path="cv_policies", iterations=it, spp=spp, seed=9, rr=rr, alpha=alpha,
                      bins=bits(b), nregions=bits(rec["nregions"]), chosen=bits(rec["chosen"]), samples=bits(rec["samples"])))
    # integrator_adaptive_tolerance: bins + number of leaves
    for integ, res, lo, hi, rule, h, tol in (("x2y2", [5], 0.0, 1.0, "simpson_trapezoidal", "default_absolute", 1e-5), ("smooth_edge2", [8, 8], 0.0, 1.0, "boole_simpson", "size_relative", 2e-4),
                                             ("ind2", [6, 5], 0.0, 1.0, "simpson_trapezoidal", "default_absolute", 2e-4), ("shade4_16", [4, 4], 0.0, 1.0, "simpson_trapezoidal", "size_relative", 4e-3),
                                             ("cubic1", [7], -0.5, 1.25, "boole_simpson", "default_relative", 1e-6), ("poly3", [3, 2], 0.1, 0.9, "simpson_trapezoidal", "size_absolute", 1e-5)):
        d = R.dim(integ)
        b, n, _ = R.adaptive_tolerance(integ, rule, h, tol, res, [lo] * d, [hi] * d)
        V.append(dict(integrand=integ, res=res, rmin=[lo] * d, rmax=[hi] * d, path="adaptive_tolerance", rule=rule, heuristic=h, tolerance=tol, size_weight=1e-5,
                      bins=bits(b), nleaves=n))
    # Fubini family: finite and infinite rests
    FUB = [("poly3", 1, [4], [0.1] * 3, [0.9] * 3), ("shade4_16", 2, [4, 3], [0.0] * 4, [1.0] * 4), ("shade5_16", 3, [2, 2], [0.0] * 5, [1.0] * 5),
           ("decay", 1, [5], [], []), ("walk", 2, [3, 2], [], []), ("walk", 2, [2, 2], [0.1, 0.2, 0.0], [0.9, 0.7, 1.0])]
    for integ, n, res, rmin, rmax in FUB:
        base = dict(integrand=integ, res=res, rmin=rmin, rmax=rmax, nfirst=n)
        V.append(dict(base, path="fubini_adaptive_mc", rule="simpson_trapezoidal", heuristic="default_absolute", iterations=12, size_weight=1e-5,
                      mc_samples=6, mc_seed=5, bins=bits(R.fubini_adaptive_mc(integ, n, "simpson_trapezoidal", "default_absolute", 12, 6, 5, res, rmin, rmax))))
        V.append(dict(base, path="fubini_mc_mc", spp=5, seed=3, mc_samples=4, mc_seed=5, bins=bits(R.fubini_mc_mc(integ, n, 5, 3, 4, 5, res, rmin, rmax))))
        if rmin and R.dim(integ) < 0:
            continue      # upstream crashes ("Empty interection", null region) for RangeInfinite with explicit non-primary entries
        V.append(dict(base, path="crespo2021_infinite", iterations=10, mc_samples=4, spp=8, seed=7,
                      bins=bits(R.crespo2021_infinite(integ, n, 10, 4, 8, 7, res, rmin, rmax))))
    # SURVEY.md §8(c) known-answer vectors (README-sized cases), kept as decimal strings the survey printed
    path = os.path.join(HERE, "reference_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(f"wrote {path}: {len(V)} vectors, {os.path.getsize(path)/1024:.0f} KiB")


if __name__ == "__main__":
    main()
