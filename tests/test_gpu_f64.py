"""-m gpu: the double-precision Newton-Cotes region family (Range<double,DIM>; north_star gate: 1e-12 relative in fp64).
Fixed rules and region->bin integration of an identical leaf table are BIT-EXACT against the oracle (itself bit-exact against the
unmodified reference instantiated with Range<double,DIM>), which is stricter than 1e-12."""
import numpy as np
import pytest
from gpu_helpers import ctx   # noqa: F401
from helpers import assert_same_bits

pytestmark = pytest.mark.gpu
DIMS = {"x2y2": 2, "ind2": 2, "cubic1": 1, "poly3": 3, "smooth_edge2": 2, "shade4_16": 4}


def _rng(integ, lo, hi):
    from viltrum_b200 import Range
    return Range([lo] * DIMS[integ], [hi] * DIMS[integ])


@pytest.mark.parametrize("integ,res,lo,hi", [("x2y2", [5], 0.0, 1.0), ("x2y2", [40, 30], 0.05, 1.1), ("smooth_edge2", [64, 64], 0.0, 1.0),
                                             ("cubic1", [300], -0.5, 1.25), ("poly3", [9, 7], 0.1, 0.9), ("shade4_16", [12, 10], 0.0, 1.0), ("ind2", [1], 0.0, 1.0)])
@pytest.mark.parametrize("rule", ["trapezoidal", "simpson", "boole"])
def test_fixed_rule_newton_cotes_fp64(ctx, port, integ, res, lo, hi, rule):
    d = DIMS[integ]
    if d >= 4 and rule == "boole":
        pytest.skip("5^4 samples: not instantiated in the oracle harness")
    init = np.linspace(-0.5, 0.5, int(np.prod(res)))
    want = port.newton_cotes_f64(integ, rule, res, [lo] * d, [hi] * d, bins=init)
    regs = ctx.regions_generate_single_f64(integ, _rng(integ, lo, hi), rule, exact=True)
    got = init.copy()
    regs.integrate_bins(got, res, _rng(integ, lo, hi))
    assert_same_bits(got, want, f"{integ} {rule} fp64")
    assert np.allclose(got, want, rtol=1e-12, atol=0)                      # the north_star gate, implied by the line above
    # the fast-math twin of the integrand stays within 1e-12 too
    fast = ctx.regions_generate_single_f64(integ, _rng(integ, lo, hi), rule, exact=False)
    g2 = init.copy(); fast.integrate_bins(g2, res, _rng(integ, lo, hi))
    added, ref = g2 - init, want - init
    assert np.allclose(added, ref, rtol=1e-12, atol=1e-12 * float(np.mean(np.abs(ref))) + 1e-300)
    regs.free(); fast.free()


@pytest.mark.parametrize("integ,res,rule,h,it", [("smooth_edge2", [64, 64], "boole_simpson", "size_relative", 3000),
                                                 ("x2y2", [17], "simpson_trapezoidal", "default_absolute", 200),
                                                 ("poly3", [10, 8], "simpson_trapezoidal", "size_relative", 300),
                                                 ("shade4_16", [16, 16], "simpson_trapezoidal", "size_relative", 400)])
def test_region_to_bin_integration_fp64_identical_leaf_table(ctx, port, integ, res, rule, h, it):
    """the reference's (oracle's) double-precision adaptive leaf table, uploaded: bins bit-identical, sharding invisible"""
    d = DIMS[integ]
    init = np.linspace(0, 1, int(np.prod(res)))
    want, reg = port.adaptive_iterations_f64(integ, rule, h, it, res, [0.0] * d, [1.0] * d, bins=init)
    regs = ctx.regions_upload_f64(rule, reg["min"], reg["max"], reg["err"], reg["dim"], reg["data"])
    back = regs.download()
    for k in ("min", "max", "err", "dim", "data"):
        assert_same_bits(back[k], reg[k], f"upload/download {k}")
    got = init.copy()
    regs.integrate_bins(got, res, _rng(integ, 0.0, 1.0))
    assert_same_bits(got, want, f"{integ} {rule} region->bin fp64")
    parts = init.copy(); nb = init.size
    regs.integrate_bins(parts, res, _rng(integ, 0.0, 1.0), shard=(0, nb // 3))
    regs.integrate_bins(parts, res, _rng(integ, 0.0, 1.0), shard=(nb // 3, nb))
    assert_same_bits(parts, want, "sharded fp64")
    regs.free()


@pytest.mark.parametrize("integ,rule,h,it,lo,hi", [("smooth_edge2", "boole_simpson", "size_relative", 3000, 0.0, 1.0), ("x2y2", "simpson_trapezoidal", "default_absolute", 200, 0.05, 1.1),
                                                   ("poly3", "simpson_trapezoidal", "size_relative", 300, 0.0, 1.0), ("shade4_16", "simpson_trapezoidal", "size_relative", 400, 0.0, 1.0),
                                                   ("ind2", "boole_simpson", "default_relative", 500, 0.0, 1.0), ("cubic1", "boole_simpson", "size_absolute", 300, -0.5, 1.25),
                                                   ("smooth_edge2", "simpson_trapezoidal", "mixed_relative_absolute", 500, 0.0, 1.0), ("x2y2", "boole_simpson", "size_relative", 0, 0.0, 1.0)])
def test_greedy_refinement_fp64_reproduces_the_reference_subdivision(ctx, port, integ, rule, h, it, lo, hi):
    """Range<double,DIM> through regions-generator-adaptive-heap.h:18-45: every sample, error and heap key a double (the greedy kernel with
    T = double, 16-byte heap entries) — region list (ranges, samples, errors, split dimensions, order) and bins bit-identical to the oracle,
    which is bit-exact against the unmodified reference instantiated with Range<double,DIM>; hence inside north_star's 1e-12 gate."""
    d = DIMS[integ]
    res = [6] * min(d, 2)
    init = np.linspace(0, 1, int(np.prod(res)))
    mixed = None
    if h.startswith("mixed"):
        mixed = dict(dimension=1, bins_weight=1.5, size_threshold_bins=1.0 / 32, size_threshold_rest=1.0 / 8, error_increase_factor=100.0)
        port.set_mixed(**mixed)
    want, reg = port.adaptive_iterations_f64(integ, rule, h, it, res, [lo] * d, [hi] * d, size_weight=1e-4, bins=init)
    parts = h.split("_")
    regs = ctx.regions_generate_adaptive_f64(integ, _rng(integ, lo, hi), rule, parts[0], parts[1], it, 1e-4, exact=True,
                                             mixed=dict(mixed, metric_rest=parts[2]) if mixed else None)
    assert len(regs) == it + 1
    got = regs.download()
    for k in ("min", "max", "err", "dim", "data"):
        assert_same_bits(got[k], reg[k], f"{integ} {rule} {h} fp64 region {k}")
    bins = init.copy()
    regs.integrate_bins(bins, res, _rng(integ, lo, hi))
    assert_same_bits(bins, want, "fp64 bins")
    assert np.allclose(bins, want, rtol=1e-12, atol=0)
    regs.free()


def test_type_mismatch_is_rejected(ctx):
    from viltrum_b200 import Vb200Error, Range
    regs = ctx.regions_generate_single_f64("x2y2", Range([0, 0], [1, 1]), "simpson")
    with pytest.raises(Vb200Error):
        ctx.check(ctx._L.vb200_regions_download(ctx._h, regs._h, None, None, None, None, None))      # float download of a double table
    regs.free()
