"""Helpers for the -m gpu parity tests (test infrastructure)."""
import numpy as np
import pytest


def z_scores(gpu, ref, var_gpu, var_ref):
    """per-bin z = (gpu-ref)/sqrt(var_gpu+var_ref) over the bins that have any variance"""
    s = np.sqrt(var_gpu + var_ref)
    ok = s > 0
    z = (gpu[ok].astype(np.float64) - ref[ok].astype(np.float64)) / s[ok]
    return z, ok


def assert_statistically_equal(gpu, ref, var_gpu, var_ref, what=""):
    """The '3 sigma' gate of BASELINE.json north_star, made a proper test over many bins: per-bin |z| <= 3 for all but the
    fraction a normal law allows (0.27% expected; 1% tolerated), no gross outlier, and — so a systematic bias cannot hide
    inside per-bin noise (SURVEY.md §8d 'Parity gates') — the z histogram is centred with unit variance."""
    z, ok = z_scores(gpu, ref, var_gpu, var_ref)
    n = z.size
    assert n > 0
    # bins with zero variance on both sides must agree to rounding
    same = ~ok
    if same.any():
        assert np.allclose(gpu[same], ref[same], rtol=1e-5, atol=1e-6), f"{what}: deterministic bins differ"
    frac = float(np.mean(np.abs(z) > 3.0))
    assert frac < 0.01 + 3.0 / n, f"{what}: {frac:.4f} of bins beyond 3 sigma"
    assert float(np.max(np.abs(z))) < 7.0, f"{what}: max |z| = {np.max(np.abs(z)):.2f}"
    if n >= 256:
        assert abs(float(np.mean(z))) < 5.0 / np.sqrt(n), f"{what}: mean z = {np.mean(z):.4f} (bias)"
        m2 = float(np.mean(z * z))
        assert 0.75 < m2 < 1.3, f"{what}: mean z^2 = {m2:.3f}"


def mc_variance(sum_f, sum_f2, n, scale):
    """variance of scale * mean(f) estimated from n samples: scale^2 * (E f^2 - (E f)^2)/n"""
    sum_f = np.asarray(sum_f, np.float64); sum_f2 = np.asarray(sum_f2, np.float64)
    var = np.maximum(sum_f2 / n - (sum_f / n) ** 2, 0.0) * n / max(n - 1, 1)
    return (scale ** 2) * var / n


@pytest.fixture(scope="session")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from viltrum_b200 import Context
    return Context(0)
