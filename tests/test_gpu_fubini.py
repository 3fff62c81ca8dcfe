"""-m gpu: the Fubini family against the oracle — integrator_fubini<N>(first, monte_carlo(m)) and
integrator_crespo2021_infinite<N> over finite and infinite rests (reference src/combination/fubini.h:51-101,
src/combination/regions-generator-fubini.h:7-28, src/control-variates/integrator-crespo2021.h:24-44).

The reference threads one mt19937 through every evaluation of g(x) = MC estimate of the integral of f(x, .); the GPU keys a
Philox stream by the evaluation point.  Streams differ, estimators are the same: parity is statistical — the mean over K
independent seeds of both implementations agrees bin by bin within 3 sigma of the standard errors (z histogram gate of
gpu_helpers), and the two estimators have the same variance."""
import numpy as np
import pytest
from gpu_helpers import ctx, assert_statistically_equal   # noqa: F401

pytestmark = pytest.mark.gpu

FINITE_DIM = {"poly3": 3, "shade4_16": 4, "shade4_64": 4, "shade5_16": 5}


def _full_range(integ, lo=0.0, hi=1.0):
    from viltrum_b200 import Range, RangeInfinite
    return Range([lo] * FINITE_DIM[integ], [hi] * FINITE_DIM[integ]) if integ in FINITE_DIM else RangeInfinite()


def _k_seeds(K, ref_fn, gpu_fn):
    refs = np.stack([ref_fn(s).astype(np.float64) for s in range(K)])
    gpus = np.stack([gpu_fn(s).astype(np.float64) for s in range(K)])
    return refs, gpus


def _gate(refs, gpus, what, var_lo=0.6, var_hi=1.6):
    K = refs.shape[0]
    if "decay" in what:      # decay's terms grow like prod 2x (heavy tail): the sample variance of 24 seeds is itself very noisy — only its scale is checked
        var_lo, var_hi = min(var_lo, 0.25), max(var_hi, 4.0)
    assert_statistically_equal(gpus.mean(axis=0), refs.mean(axis=0), gpus.var(axis=0, ddof=1) / K, refs.var(axis=0, ddof=1) / K, what)
    ratio = (gpus.var(axis=0, ddof=1).mean() + 1e-30) / (refs.var(axis=0, ddof=1).mean() + 1e-30)
    assert var_lo < ratio < var_hi, f"{what}: variance ratio {ratio:.3f}"


@pytest.mark.parametrize("integ,n,res", [("poly3", 1, [24]), ("poly3", 2, [8, 6]), ("shade4_16", 2, [12, 12]), ("shade5_16", 3, [8, 8]),
                                          ("decay", 1, [16]), ("walk", 2, [8, 8]), ("walk", 1, [16])])
def test_fubini_per_bin_mc_first(ctx, port, integ, n, res):
    """integrator_fubini<N>(monte_carlo_per_bin_parallel(spp), monte_carlo(m)): K1 with the adapter as its integrand ('+=')"""
    from viltrum_b200 import integrate, integrator_fubini, monte_carlo_per_bin_parallel, monte_carlo
    nb = int(np.prod(res)); spp, m, K = 16, 4, 24
    rng = _full_range(integ)

    def gpu(s):
        b = np.zeros(nb, np.float32)
        integrate(integrator_fubini(n, monte_carlo_per_bin_parallel(spp, seed=s), monte_carlo(m, seed=1000 + s)), b, res, integ, rng, ctx=ctx)
        return b
    refs, gpus = _k_seeds(K, lambda s: port.fubini_mc_mc(integ, n, spp, 50 + s, m, 70 + s, res, rng.min, rng.max), gpu)
    _gate(refs, gpus, f"fubini<{n}> mc/mc {integ}")
    # '+=' semantics: the estimate is added to what the bins held
    b = np.full(nb, 2.0, np.float32)
    integrate(integrator_fubini(n, monte_carlo_per_bin_parallel(spp, seed=0), monte_carlo(m, seed=1000)), b, res, integ, rng, ctx=ctx)
    assert np.allclose(b - 2.0, gpus[0], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("integ,n,res,it", [("poly3", 1, [16], 24), ("shade4_16", 2, [8, 8], 64), ("decay", 1, [12], 24), ("walk", 2, [6, 6], 48)])
@pytest.mark.parametrize("batch", [1, 0])
def test_fubini_adaptive_first(ctx, port, integ, n, res, it, batch):
    """integrator_fubini<N>(integrator_adaptive_iterations(...), monte_carlo(m)): the greedy generator over a noisy integrand.
    batch=1 is the reference's greedy order (persistent kernel calling the adapter), batch=0 the batched generator through the
    eval thunk; both must be statistically the reference's estimator."""
    from viltrum_b200 import (integrate, integrator_fubini, integrator_adaptive_iterations, monte_carlo, nested, error_heuristic_default,
                              error_metric_absolute)
    nb = int(np.prod(res)); m, K = 16, 24
    rng = _full_range(integ)

    def gpu(s):
        b = np.zeros(nb, np.float32)
        first = integrator_adaptive_iterations(nested("simpson", "trapezoidal"), error_heuristic_default(error_metric_absolute()), it, batch=batch)
        integrate(integrator_fubini(n, first, monte_carlo(m, seed=300 + s)), b, res, integ, rng, ctx=ctx)
        return b
    refs, gpus = _k_seeds(K, lambda s: port.fubini_adaptive_mc(integ, n, "simpson_trapezoidal", "default_absolute", it, m, 90 + s, res, rng.min, rng.max), gpu)
    # batched refinement picks a different (equally valid) set of regions: same expectation, somewhat different noise
    _gate(refs, gpus, f"fubini<{n}> adaptive {integ} batch={batch}", 0.4 if batch == 0 else 0.6, 2.5 if batch == 0 else 1.6)


@pytest.mark.parametrize("integ,n,res,it,mc,spp", [("shade4_16", 2, [12, 12], 64, 4, 16), ("shade5_16", 3, [6, 6], 48, 4, 16), ("poly3", 1, [16], 16, 4, 16),
                                                   ("walk", 2, [8, 8], 48, 8, 32), ("decay", 1, [12], 16, 8, 32), ("decay", 2, [6, 4], 32, 8, 32)])
def test_crespo2021_infinite(ctx, port, integ, n, res, it, mc, spp):
    """integrator_crespo2021_infinite<N>: region table of the noisy g, control-variate integral + residual samples of f itself"""
    from viltrum_b200 import integrate, integrator_crespo2021_infinite
    nb = int(np.prod(res)); K = 24
    rng = _full_range(integ)

    def gpu(s):
        b = np.full(nb, -5.0, np.float32)          # '=': previous contents must not matter
        integrate(integrator_crespo2021_infinite(n, it, mc, spp, seed=s), b, res, integ, rng, ctx=ctx)
        return b
    refs, gpus = _k_seeds(K, lambda s: port.crespo2021_infinite(integ, n, it, mc, spp, 200 + s, res, rng.min, rng.max), gpu)
    _gate(refs, gpus, f"crespo2021_infinite<{n}> {integ}")


def test_fubini_adapter_is_an_ordinary_integrand(ctx):
    """the adapter goes wherever an integrand goes: resident bins, shards, the global scatter sampler; unknown combinations fail loudly"""
    import torch
    from viltrum_b200 import FubiniIntegrand, Range, range_split_at, _capi
    full = Range([0.0] * 4, [1.0] * 4)
    first, rest = range_split_at(2, full)
    g = FubiniIntegrand(ctx, "shade4_64", 2, rest, 8, 5)
    res = [64, 48]; nb = res[0] * res[1]
    whole = torch.zeros(nb, dtype=torch.float32, device="cuda")
    ctx.mc_per_bin(g, whole, res, first, 32, 9)
    parts = torch.zeros(nb, dtype=torch.float32, device="cuda")
    ctx.mc_per_bin(g, parts, res, first, 32, 9, shard=(0, 1000)); ctx.mc_per_bin(g, parts, res, first, 32, 9, shard=(1000, nb))
    ctx.synchronize()
    assert torch.equal(whole, parts)
    assert abs(float(whole.mean()) - 0.14326) < 3e-3                      # integral of shade4<64>
    sc = np.zeros(nb, np.float32)
    ctx.monte_carlo(g, sc, res, first, 200000, 3)
    assert abs(float(sc.mean()) - 0.14326) < 3e-3
    g.free()
    with pytest.raises(KeyError):
        FubiniIntegrand(ctx, "shade4_64", 3, Range([0.0], [1.0]), 8, 5)
    with pytest.raises(KeyError):
        FubiniIntegrand(ctx, "shade4_64", 2, Range([0.0], [1.0]), 8, 5)      # wrong number of rest entries
