"""-m gpu: BASELINE.json's FULL sizes for configs 3, 4 and 5 through size-independent properties (the oracle cannot finish these in
seconds: C4 alone is ~15 minutes of CPU): mean of bins = the integral, leaves tile the domain, region counts per bin match the
survey's measurement of the reference, shards reproduce the whole, write semantics."""
import numpy as np
import pytest
from gpu_helpers import ctx   # noqa: F401
from helpers import assert_same_bits

pytestmark = pytest.mark.gpu


def test_config3_full_size_batched(ctx):
    """512x512 bins, nested(boole,simpson), size/relative 1e-5, 10^6 splits (batched refinement) on smooth_edge2"""
    import torch
    from viltrum_b200 import Range
    rng = Range([0, 0], [1, 1])
    regs = ctx.regions_generate_adaptive("smooth_edge2", rng, "boole_simpson", "size", "relative", 1000000, 1e-5, batch=0, exact=True)
    assert len(regs) == 1000001
    bins = torch.zeros(512 * 512, dtype=torch.float32, device="cuda")
    regs.integrate_bins(bins, [512, 512], rng)
    ctx.synchronize()
    b = bins.cpu().numpy()
    analytic = 0.5 + 2 / 9 - 2 / 45 + 0.75 * np.pi * 0.09
    assert abs(float(b.mean(dtype=np.float64)) - analytic) < 2e-6            # the reference's own 10^6-iteration mean is 0.889836
    t = regs.download()
    vol = np.prod(t["max"].astype(np.float64) - t["min"].astype(np.float64), axis=1)
    assert abs(vol.sum() - 1.0) < 1e-6 and vol.min() > 0
    twice = bins.clone()
    regs.integrate_bins(twice, [512, 512], rng); ctx.synchronize()
    assert np.allclose(twice.cpu().numpy(), 2 * b, rtol=1e-6)                # '+='
    regs.free()


def test_config4_full_size(ctx):
    """1024x1024 bins, crespo2021(65536 iterations, 64 spp) on shade5<64>: SURVEY.md measured 967.6 regions per bin and
    1.015e9 (bin, region) pairs for the reference's own (greedy) subdivision — reproduced here with the exact generator"""
    import torch
    from viltrum_b200 import Range
    rng = Range([0] * 5, [1] * 5)
    regs = ctx.regions_generate_adaptive("shade5_64", rng, "simpson_trapezoidal", "size", "relative", 65536, 1e-5, batch=1, exact=True)
    nb = 1 << 20
    bins = torch.full((nb,), 123.0, dtype=torch.float32, device="cuda")
    nreg = torch.zeros(nb, dtype=torch.int32, device="cuda")
    approx = torch.zeros(nb, dtype=torch.float32, device="cuda")
    regs.cv_integrate("shade5_64", bins, [1024, 1024], rng, 64, 0, nregions=nreg, approx=approx)
    ctx.synchronize()
    assert abs(float(nreg.double().mean()) - 967.6) < 0.1 and abs(float(nreg.double().sum()) - 1.015e9) < 2e6
    b = bins.cpu().numpy()
    assert abs(float(b.mean(dtype=np.float64)) - 0.14326) < 1.5e-4           # '=' : the 123.0 fill is gone
    assert abs(float(approx.double().mean()) - 0.14326) < 3e-3               # the control variate alone is already close
    # a slab of rows computed alone is bit-identical to the same rows of the whole (multi-GPU sharding)
    part = torch.zeros(nb, dtype=torch.float32, device="cuda")
    regs.cv_integrate("shade5_64", part, [1024, 1024], rng, 64, 0, shard=(300 * 1024, 364 * 1024))
    ctx.synchronize()
    assert_same_bits(part.cpu().numpy()[300 * 1024:364 * 1024], b[300 * 1024:364 * 1024], "C4 slab")
    regs.free()


def test_config5_full_size(ctx):
    """2048x2048 bins, 256 spp random walk over range_primary_infinite: mean 1.0133; '+=' doubles"""
    import torch
    from viltrum_b200 import RangeInfinite
    d = torch.zeros(1 << 22, dtype=torch.float32, device="cuda")
    ctx.mc_per_bin_inf("walk", d, [2048, 2048], RangeInfinite(), 256, 0); ctx.synchronize()
    a = d.cpu().numpy().copy()
    assert abs(float(a.mean(dtype=np.float64)) - 1.0133) < 5e-4
    ctx.mc_per_bin_inf("walk", d, [2048, 2048], RangeInfinite(), 256, 0); ctx.synchronize()
    assert_same_bits(d.cpu().numpy(), (a.astype(np.float64) * 2).astype(np.float32), "second += doubles")
