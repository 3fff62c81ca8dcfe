"""-m gpu: integrator_adaptive_tolerance (reference src/nested/integrator-adaptive-tolerance.h:15-39).
The reference recurses depth first; the device splits every failing region of a round at once and sorts the leaves by their
root-to-leaf path.  The result must be the reference's own leaf list IN ITS ORDER (ranges, samples, err, dim — bit for bit, against
the oracle port, which is itself pinned to the unmodified reference's bins and leaf counts) and therefore its bins, bit for bit."""
import numpy as np
import pytest
from gpu_helpers import ctx   # noqa: F401
from helpers import load_golden, f32, assert_same_bits

pytestmark = pytest.mark.gpu
DIMS = {"x2y2": 2, "ind2": 2, "cubic1": 1, "poly3": 3, "shade4_16": 4, "shade4_64": 4, "shade5_16": 5, "smooth_edge2": 2}


def _rng(integ, lo=0.0, hi=1.0):
    from viltrum_b200 import Range
    return Range([lo] * DIMS[integ], [hi] * DIMS[integ])


CASES = [("x2y2", [17], "simpson_trapezoidal", "default", "absolute", 1e-6), ("smooth_edge2", [64, 64], "boole_simpson", "default", "absolute", 1e-7),
         ("ind2", [24, 20], "simpson_trapezoidal", "default", "absolute", 2e-5), ("poly3", [10, 8], "simpson_trapezoidal", "size", "absolute", 2e-6),
         ("shade4_16", [16, 16], "simpson_trapezoidal", "size", "relative", 1.5e-3), ("shade4_16", [8, 8], "simpson_trapezoidal", "default", "absolute", 1e-4), ("shade5_16", [8, 8], "simpson_trapezoidal", "default", "absolute", 3e-4),
         ("cubic1", [64], "simpson_trapezoidal", "default", "absolute", 1e-7), ("smooth_edge2", [16, 16], "simpson_trapezoidal", "default", "relative", 1e-4),
         ("x2y2", [4, 4], "boole_simpson", "size", "relative", 10.0)]      # tolerance above the root's error: one leaf


@pytest.mark.parametrize("integ,res,rule,hk,mk,tol", CASES)
def test_leaf_list_in_depth_first_order_and_bins_bit_exact(ctx, port, integ, res, rule, hk, mk, tol):
    from viltrum_b200 import integrate, integrator_adaptive_tolerance, nested, error_heuristic_default, error_heuristic_size, error_metric_absolute, error_metric_relative
    d = DIMS[integ]
    init = np.linspace(-1, 1, int(np.prod(res))).astype(np.float32)
    want, n, reg = port.adaptive_tolerance(integ, rule, f"{hk}_{mk}", tol, res, [0.0] * d, [1.0] * d, bins=init, reg_cap=2_000_000)
    regs = ctx.regions_generate_tolerance(integ, _rng(integ), rule, hk, mk, tol, 1e-5, exact=True)
    assert len(regs) == n
    got = regs.download()
    for k in ("min", "max", "err", "dim", "data"):
        assert_same_bits(got[k], reg[k], f"{integ} leaf {k} ({n} leaves)")
    regs.free()
    metric = error_metric_absolute() if mk == "absolute" else error_metric_relative()
    eh = error_heuristic_default(metric) if hk == "default" else error_heuristic_size(metric, 1e-5)
    bins = init.copy()
    h, l = rule.split("_")
    integrate(integrator_adaptive_tolerance(nested(h, l), eh, tol), bins, res, integ, _rng(integ), ctx=ctx)
    assert_same_bits(bins, want, f"{integ} bins")


def test_golden_reference_vectors(ctx):
    from viltrum_b200 import Range
    n = 0
    for v in load_golden():
        if v["path"] != "adaptive_tolerance":
            continue
        hk, mk = v["heuristic"].split("_")
        regs = ctx.regions_generate_tolerance(v["integrand"], Range(v["rmin"], v["rmax"]), v["rule"], hk, mk, v["tolerance"], v["size_weight"], exact=True)
        assert len(regs) == v["nleaves"]
        bins = np.zeros(int(np.prod(v["res"])), np.float32)
        regs.integrate_bins(bins, v["res"], Range(v["rmin"], v["rmax"]))
        assert_same_bits(bins, f32(v["bins"]), f"{v['integrand']} vs the unmodified reference")
        regs.free(); n += 1
    assert n == 6


def test_many_leaves_growth_and_limits(ctx, port):
    """a table that outgrows its initial capacity several times; the region cap and the depth cap fail loudly"""
    from viltrum_b200 import Vb200Error
    want, n, reg = port.adaptive_tolerance("smooth_edge2", "simpson_trapezoidal", "default_absolute", 1e-8, [32, 32], [0, 0], [1, 1], reg_cap=2_000_000)
    assert n > 20000
    regs = ctx.regions_generate_tolerance("smooth_edge2", _rng("smooth_edge2"), "simpson_trapezoidal", "default", "absolute", 1e-8, 1e-5, exact=True)
    assert len(regs) == n
    got = regs.download()
    assert_same_bits(got["min"], reg["min"], "leaf order"); assert_same_bits(got["data"], reg["data"], "leaf samples")
    bins = np.zeros(32 * 32, np.float32)
    regs.integrate_bins(bins, [32, 32], _rng("smooth_edge2"))
    assert_same_bits(bins, want, "bins")
    regs.free()
    with pytest.raises(Vb200Error):
        ctx.regions_generate_tolerance("smooth_edge2", _rng("smooth_edge2"), "simpson_trapezoidal", "default", "absolute", 1e-8, 1e-5, max_regions=1000)
    with pytest.raises(Vb200Error):      # ind2's discontinuity never gets below an absurd tolerance: the reference would recurse forever
        ctx.regions_generate_tolerance("ind2", _rng("ind2"), "simpson_trapezoidal", "default", "relative", 1e-30, 1e-5, max_regions=1 << 22)
