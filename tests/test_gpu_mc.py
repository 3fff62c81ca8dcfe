"""-m gpu: per-bin Monte Carlo through the C ABI against the oracle.
  * statistical parity (3 sigma gate, z histogram) on seeded inputs the oracle finishes in seconds,
  * bit-exact parity in sample-replay mode (the oracle's own sample points),
  * committed golden vectors (samples recorded from the unmodified reference),
  * size-independent properties at BASELINE.json's full size."""
import ctypes
import numpy as np
import pytest
from gpu_helpers import ctx, assert_statistically_equal, mc_variance   # noqa: F401
from helpers import load_golden, f32, assert_same_bits

pytestmark = pytest.mark.gpu


def _rng(vb, integ, lo=0.0, hi=1.0):
    from viltrum_b200 import Range
    d = {"x2y2": 2, "ind2": 2, "cubic1": 1, "poly3": 3, "shade4_16": 4, "shade4_64": 4, "shade5_16": 5, "shade5_64": 5, "smooth_edge2": 2}[integ]
    return Range([lo] * d, [hi] * d)


def test_counter_layout_and_affine_map(ctx):
    """spp=1: the single sample of bin b is Philox(key=seed, ctr=(b,0,0,0)) mapped into the bin box — predicted on the host."""
    from viltrum_b200 import _capi, Range
    L = _capi.lib()
    res, seed = [7, 5], 0x1234567890ABCDEF
    bins = np.zeros(35, np.float32)
    ctx.mc_per_bin("x2y2", bins, res, Range([0.25, -1.0], [2.0, 3.0]), 1, seed, generator="philox")
    want = np.zeros(35, np.float32)
    for b in range(35):
        c = (ctypes.c_uint32 * 4)(b, 0, 0, 0); k = (ctypes.c_uint32 * 2)(seed & 0xffffffff, seed >> 32); o = (ctypes.c_uint32 * 4)()
        L.vb200_philox4x32_10(c, k, o)
        u = [np.float32(o[i] >> 8) * np.float32(2.0 ** -24) for i in range(2)]
        p = [b % 7, b // 7]
        x = []
        for i, (lo, hi, r) in enumerate(((0.25, 2.0, 7), (-1.0, 3.0, 5))):
            dr = np.float32(np.float32(hi - lo) / np.float32(r))
            a = np.float32(lo) + np.float32(p[i]) * dr; bb = np.float32(lo) + np.float32(p[i] + 1) * dr
            x.append(np.float64(u[i]) * np.float64(np.float32(bb - a)) + np.float64(a))     # fma: single rounding
        x = [np.float32(v) for v in x]
        f = np.float32(x[0] * x[0]) + np.float32(x[1] * x[1])
        vol = np.float32(np.float32(2.0 - 0.25) * np.float32(3.0 + 1.0))
        want[b] = np.float32(np.float64(f) * (np.float64(vol) / 1.0))
    assert np.allclose(bins, want, rtol=3e-7, atol=0)


def test_counter_layout_narrow_fields(ctx):
    """Grids with >= 256 bins along every binned dimension draw the in-bin coordinates as 16-bit fields (mc_per_bin.cuh GroupDraws):
    spp=1 -> sample 0 of bin b takes the low halves of words 0 and 1 of Philox(key=seed, ctr=(b,0,0,0)) — predicted on the host."""
    from viltrum_b200 import _capi, Range
    L = _capi.lib()
    res, seed = [256, 300], 0x0FEDCBA987654321
    nb = res[0] * res[1]
    bins = np.zeros(nb, np.float32)
    ctx.mc_per_bin("x2y2", bins, res, Range([0.25, -1.0], [2.0, 3.0]), 1, seed, generator="philox")
    vol = np.float32(np.float32(2.0 - 0.25) * np.float32(3.0 + 1.0))
    for b in range(0, nb, 97):
        c = (ctypes.c_uint32 * 4)(b, 0, 0, 0); k = (ctypes.c_uint32 * 2)(seed & 0xffffffff, seed >> 32); o = (ctypes.c_uint32 * 4)()
        L.vb200_philox4x32_10(c, k, o)
        p = [b % res[0], b // res[0]]
        x = []
        for i, (lo, hi, r) in enumerate(((0.25, 2.0, res[0]), (-1.0, 3.0, res[1]))):
            dr = np.float32(np.float32(hi - lo) / np.float32(r))
            a = np.float32(lo) + np.float32(p[i]) * dr; bb = np.float32(lo) + np.float32(p[i] + 1) * dr
            x.append(np.float32(np.float64(o[i] & 0xffff) * 2.0 ** -16 * np.float64(np.float32(bb - a)) + np.float64(a)))
            assert a <= x[i] <= bb
        f = np.float64(x[0]) ** 2 + np.float64(x[1]) ** 2
        want = f * np.float64(vol)
        # the kernel folds the fields' 2^23 offset into the bin corner: the sample may sit a fraction of a lattice step away
        assert abs(bins[b] - want) <= 2e-6 * abs(want), (b, bins[b], want)


@pytest.mark.parametrize("lattice24", [False, True])
def test_stream_layout_xoshiro(ctx, lattice24):
    """Default generator: one xoshiro128++ stream per (bin, lane sub-stream), state = Philox4x32-10(key=seed, ctr=(bin lo, bin hi, sub, 'strm')).
    spp=1 -> the sample of bin b is cut from the first words of sub-stream 0 (16-bit fields on this >= 256-bin grid, 24-bit fields with
    VB200_MC_LATTICE24) — predicted on the host through vb200_philox4x32_10 + vb200_xoshiro128pp."""
    from viltrum_b200 import _capi, Range
    L = _capi.lib()
    res, seed = [256, 300], 0x0FEDCBA987654321
    nb = res[0] * res[1]
    bins = np.zeros(nb, np.float32)
    ctx.mc_per_bin("x2y2", bins, res, Range([0.25, -1.0], [2.0, 3.0]), 1, seed, lattice24=lattice24)
    vol = np.float32(np.float32(2.0 - 0.25) * np.float32(3.0 + 1.0))
    for b in range(0, nb, 97):
        c = (ctypes.c_uint32 * 4)(b, 0, 0, 0x7374726d); k = (ctypes.c_uint32 * 2)(seed & 0xffffffff, seed >> 32); st = (ctypes.c_uint32 * 4)()
        L.vb200_philox4x32_10(c, k, st)
        o = (ctypes.c_uint32 * 2)()
        L.vb200_xoshiro128pp(st, 2, o)
        p = [b % res[0], b // res[0]]
        x = []
        for i, (lo, hi, r) in enumerate(((0.25, 2.0, res[0]), (-1.0, 3.0, res[1]))):
            dr = np.float32(np.float32(hi - lo) / np.float32(r))
            a = np.float32(lo) + np.float32(p[i]) * dr; bb = np.float32(lo) + np.float32(p[i] + 1) * dr
            u = np.float64(o[i] >> 8) * 2.0 ** -24 if lattice24 else np.float64(o[i] & 0xffff) * 2.0 ** -16
            x.append(np.float32(u * np.float64(np.float32(bb - a)) + np.float64(a)))
        want = (np.float64(x[0]) ** 2 + np.float64(x[1]) ** 2) * np.float64(vol)
        assert abs(bins[b] - want) <= 2e-6 * abs(want), (b, bins[b], want)


@pytest.mark.parametrize("generator", ["xoshiro", "philox"])
@pytest.mark.parametrize("integ,res,spp", [("shade4_64", [256, 256], 64), ("shade4_16", [64, 48], 37), ("x2y2", [100], 256),
                                           ("poly3", [12, 10, 6], 32), ("shade5_16", [32, 32], 16), ("ind2", [24, 24], 128)])
@pytest.mark.parametrize("flavor", ["mc_per_bin_parallel", "per_bin_parallel_mc"])
def test_statistical_parity(ctx, port, integ, res, spp, flavor, generator):
    from viltrum_b200 import _capi
    rng = _rng(None, integ, 0.0, 1.0)
    nb = int(np.prod(res))
    g = np.zeros(nb, np.float32); s1 = np.zeros(nb, np.float32); s2 = np.zeros(nb, np.float32)
    fl = _capi.MC_PER_BIN if flavor == "mc_per_bin_parallel" else _capi.PER_BIN_MC
    ctx.mc_per_bin(integ, g, res, rng, spp, 1234, fl, sum_f=s1, sum_f2=s2, generator=generator)
    if len(res) <= 2:
        r, _, r1, r2 = getattr(port, flavor)(integ, res, rng.min, rng.max, spp, 99, record=True)
    else:   # the oracle bins over <= 2 dims: compare per-bin against an independent GPU seed instead, and the mean against the oracle
        r = np.zeros(nb, np.float32); r1 = np.zeros(nb, np.float32); r2 = np.zeros(nb, np.float32)
        ctx.mc_per_bin(integ, r, res, rng, spp, 4321, fl, sum_f=r1, sum_f2=r2)
        o = port.mc_per_bin_parallel(integ, res[:2], rng.min, rng.max, 4096, 5)
        assert abs(float(np.mean(g)) - float(np.mean(o))) < 0.01
    vol = float(np.prod(np.asarray(rng.max) - np.asarray(rng.min)))
    assert_statistically_equal(g, r, mc_variance(s1, s2, spp, vol), mc_variance(r1, r2, spp, vol), f"{integ} {flavor}")
    # the kernel's own scaling: bins == sum_f * vol/spp.  The per_bin_parallel(monte_carlo) flavor scales by the bin box's own
    # float volume, whose extents (differences of numbers near 1) carry ~1e-7 absolute error: a few 1e-5 relative at 100 bins/dim.
    assert np.allclose(g, s1.astype(np.float64) * vol / spp, rtol=(2e-6 if flavor == "mc_per_bin_parallel" else 1e-4), atol=1e-7)


@pytest.mark.parametrize("integ,res,spp,lo,hi", [("shade4_64", [16, 12], 64, 0.0, 1.0), ("shade4_16", [9, 7], 5, 0.05, 1.1), ("x2y2", [33], 20, -0.5, 1.5),
                                                 ("cubic1", [17], 9, 0.0, 2.0), ("shade5_64", [6, 5], 16, 0.0, 1.0), ("smooth_edge2", [20, 20], 8, 0.0, 1.0),
                                                 ("poly3", [5, 4], 11, 0.1, 0.9), ("ind2", [1], 40, 0.0, 1.0)])
@pytest.mark.parametrize("flavor", ["mc_per_bin_parallel", "per_bin_parallel_mc"])
def test_replay_bit_exact(ctx, port, integ, res, spp, lo, hi, flavor):
    """north_star: 'match exactly in a sample-replay mode that feeds the reference's sample points'."""
    from viltrum_b200 import _capi
    rng = _rng(None, integ, lo, hi)
    nb = int(np.prod(res))
    init = np.linspace(-1, 1, nb).astype(np.float32)
    want, samples, _, _ = getattr(port, flavor)(integ, res, rng.min, rng.max, spp, 7, bins=init, record=True)
    got = init.copy()
    fl = _capi.MC_PER_BIN if flavor == "mc_per_bin_parallel" else _capi.PER_BIN_MC
    ctx.mc_per_bin_replay(integ, got, res, rng, spp, np.ascontiguousarray(samples), fl)
    assert_same_bits(got, want, f"{integ} {flavor}")


def test_replay_golden_reference_vectors(ctx):
    """samples and bins recorded from the UNMODIFIED reference (tests/golden/reference_vectors.json)"""
    from viltrum_b200 import _capi, Range
    n = 0
    for v in load_golden():
        if v["path"] not in ("mc_per_bin_parallel", "per_bin_parallel_mc"):
            continue
        rng = Range(v["rmin"], v["rmax"])
        nb = int(np.prod(v["res"]))
        got = np.zeros(nb, np.float32)
        fl = _capi.MC_PER_BIN if v["path"] == "mc_per_bin_parallel" else _capi.PER_BIN_MC
        ctx.mc_per_bin_replay(v["integrand"], got, v["res"], rng, v["spp"], np.ascontiguousarray(f32(v["samples"])), fl)
        assert_same_bits(got, f32(v["bins"]), f"{v['path']} {v['integrand']} {v['res']}")
        n += 1
    assert n >= 20


@pytest.mark.parametrize("generator", ["xoshiro", "philox"])
def test_sharding_is_invisible(ctx, generator):
    """Streams / counters are keyed by the global bin index: any split of the grid gives the same bits."""
    res, spp = [50, 30], 48
    rng = _rng(None, "shade4_16")
    full = np.zeros(1500, np.float32)
    ctx.mc_per_bin("shade4_16", full, res, rng, spp, 5, generator=generator)
    parts = np.zeros(1500, np.float32)
    for lo, hi in ((0, 1), (1, 700), (700, 701), (701, 1500)):
        ctx.mc_per_bin("shade4_16", parts, res, rng, spp, 5, shard=(lo, hi), generator=generator)
    assert_same_bits(full, parts, "sharded vs whole")
    other = np.zeros(1500, np.float32)
    ctx.mc_per_bin("shade4_16", other, res, rng, spp, 6, generator=generator)
    assert not np.array_equal(full, other)


def test_generators_are_distinct_streams_of_the_same_estimator(ctx):
    """the two generators must not share samples, and option bits outside the ABI are rejected"""
    from viltrum_b200 import Vb200Error, _capi
    res, spp = [64, 64], 64
    rng = _rng(None, "shade4_16")
    a = np.zeros(4096, np.float32); b = np.zeros(4096, np.float32)
    ctx.mc_per_bin("shade4_16", a, res, rng, spp, 5, generator="xoshiro")
    ctx.mc_per_bin("shade4_16", b, res, rng, spp, 5, generator="philox")
    assert not np.array_equal(a, b)
    assert abs(float(a.mean()) - float(b.mean())) < 5e-3
    p = ctx._mc_params(4, res, rng, spp, 5, _capi.MC_PER_BIN, None, options=8)
    import ctypes as ct
    assert ctx._L.vb200_mc_per_bin(ctx._h, ctx.integrand("shade4_16"), ct.byref(p), a.ctypes.data, _capi.HOST, None, None) == -2


def test_write_semantics_and_device_path(ctx):
    """'+=' for monte_carlo_per_bin_parallel, '=' for integrator_per_bin_parallel; device-resident
    bins give the same bits as host bins."""
    import torch
    from viltrum_b200 import _capi
    res, spp = [40, 25], 32
    rng = _rng(None, "x2y2")
    base = np.zeros(1000, np.float32)
    ctx.mc_per_bin("x2y2", base, res, rng, spp, 3)
    acc = np.full(1000, 10.0, np.float32)
    ctx.mc_per_bin("x2y2", acc, res, rng, spp, 3)
    assert_same_bits(acc, (np.float64(10.0) + base.astype(np.float64)).astype(np.float32), "+= on host bins")
    over = np.full(1000, 10.0, np.float32)
    ctx.mc_per_bin("x2y2", over, res, rng, spp, 3, _capi.PER_BIN_MC)
    assert abs(float(np.mean(over)) - 2.0 / 3.0) < 0.02 and float(np.max(over)) < 3.0           # overwritten, not accumulated
    dev = torch.full((1000,), 10.0, dtype=torch.float32, device="cuda")
    ctx.mc_per_bin("x2y2", dev, res, rng, spp, 3); ctx.synchronize()
    assert_same_bits(dev.cpu().numpy(), acc, "device += vs host +=")
    dev2 = torch.full((1000,), 10.0, dtype=torch.float32, device="cuda")
    ctx.mc_per_bin("x2y2", dev2, res, rng, spp, 3, _capi.PER_BIN_MC); ctx.synchronize()
    assert_same_bits(dev2.cpu().numpy(), over, "device = vs host =")


@pytest.mark.parametrize("res,spp", [([1], 1), ([3], 2), ([257], 3), ([31, 1], 1000), ([2, 2, 2], 17), ([1, 1], 4096)])
def test_edge_shapes(ctx, port, res, spp):
    """ragged grids (fewer bins than one CTA tile, spp not a multiple of the lanes per bin, 3-D bins, a single bin)"""
    integ = "poly3"
    rng = _rng(None, integ)
    nb = int(np.prod(res))
    g = np.zeros(nb, np.float32)
    ctx.mc_per_bin(integ, g, res, rng, spp, 11)
    assert np.all(np.isfinite(g))
    exact = 0.25 + 1.0 / 6.0 + 0.5        # integral of x*y + y*z^2 + 0.5 over the unit cube
    tol = 6.0 * 0.35 / np.sqrt(nb * spp) + 1e-6
    assert abs(float(np.mean(g)) - exact) < tol


def test_errors_are_reported(ctx):
    from viltrum_b200 import Vb200Error, Range
    b = np.zeros(4, np.float32)
    with pytest.raises(Vb200Error):
        ctx.mc_per_bin("x2y2", b, [4], Range([0] * 3, [1] * 3), 4, 0)             # range/integrand dimension mismatch
    with pytest.raises(Vb200Error):
        ctx.mc_per_bin("x2y2", b, [4], Range([0] * 2, [1] * 2), 0, 0)             # spp = 0
    with pytest.raises(Vb200Error):
        ctx.mc_per_bin("x2y2", b, [4], Range([0] * 2, [1] * 2), 4, 0, shard=(3, 9))
    with pytest.raises(Vb200Error):
        ctx.mc_per_bin("walk", b, [4], Range([0] * 2, [1] * 2), 4, 0)             # sequence integrand on the finite path


def test_full_size_properties(ctx):
    """BASELINE.json config 2 at full size (1024x1024 bins, 64 spp, shade4<64>): mean of bins = the integral
    (0.14326), doubling through '+=' is exact, distinct seeds decorrelate, no bin is left untouched."""
    import torch
    res, spp = [1024, 1024], 64
    rng = _rng(None, "shade4_64")
    d = torch.zeros(1 << 20, dtype=torch.float32, device="cuda")
    ctx.mc_per_bin("shade4_64", d, res, rng, spp, 0); ctx.synchronize()
    a = d.cpu().numpy().copy()
    assert abs(float(a.mean(dtype=np.float64)) - 0.14326) < 2e-4
    ctx.mc_per_bin("shade4_64", d, res, rng, spp, 0); ctx.synchronize()
    assert_same_bits(d.cpu().numpy(), (a.astype(np.float64) * 2).astype(np.float32), "second += doubles")
    h = np.zeros(1 << 20, np.float32)
    ctx.mc_per_bin("shade4_64", h, res, rng, spp, 0)
    assert_same_bits(h, a, "host path == device path at full size")
    e = np.zeros(1 << 20, np.float32)
    ctx.mc_per_bin("shade4_64", e, res, rng, spp, 1)
    assert not np.array_equal(a, e) and abs(float(e.mean(dtype=np.float64)) - 0.14326) < 2e-4   # same image, independent noise
    assert float(np.mean(a > 0)) > 0.9


_FULL_SIZE_REFERENCE = {}


def _full_size_reference(flavor):
    """the unmodified reference's own estimate AND its own per-bin moments at C2's full size, computed once per session: the
    multi-threaded build where it exists (its bins are bit-identical to the serial build's — the per-bin seeds are drawn up front,
    monte-carlo-per-bin-parallel.h:50-54 — and it books a sample's f, f^2 to the bin its coordinates fall in), else the serial one"""
    import os
    import pyoracle
    if flavor not in _FULL_SIZE_REFERENCE:
        kind = "reference-mt" if pyoracle.available("reference-mt") else "reference"
        if not pyoracle.available(kind):
            pytest.skip("oracle/_ref was not built (needs /root/reference at build time)")
        O = pyoracle.load(kind)
        O.set_threads(len(os.sched_getaffinity(0)))
        ref, _, r1, r2 = getattr(O, flavor)("shade4_64", [1024, 1024], [0.0] * 4, [1.0] * 4, 64, 2024, record="moments")
        _FULL_SIZE_REFERENCE[flavor] = (ref, r1, r2)
    return _FULL_SIZE_REFERENCE[flavor]


@pytest.mark.parametrize("flavor", ["mc_per_bin_parallel", "per_bin_parallel_mc"])
@pytest.mark.parametrize("generator", ["xoshiro", "philox"])
def test_full_size_per_bin_parity_against_the_reference(ctx, flavor, generator):
    """BASELINE config 2 at FULL size: every one of the 2^20 bins within 3 sigma of the unmodified reference's estimate of the same bin,
    sigma^2 = the reference's own estimated variance + ours (each side from its own per-bin moments: a shared variance estimate would
    correlate numerator and denominator and fatten the tails).  The grid has >= 256 bins per axis, so this is the 16-bit in-bin lattice."""
    from viltrum_b200 import _capi
    ref, r1, r2 = _full_size_reference(flavor)
    res, spp, nb = [1024, 1024], 64, 1 << 20
    rng = _rng(None, "shade4_64")
    g = np.zeros(nb, np.float32); s1 = np.zeros(nb, np.float32); s2 = np.zeros(nb, np.float32)
    fl = _capi.MC_PER_BIN if flavor == "mc_per_bin_parallel" else _capi.PER_BIN_MC
    ctx.mc_per_bin("shade4_64", g, res, rng, spp, 7, fl, sum_f=s1, sum_f2=s2, generator=generator)
    assert_statistically_equal(g, ref, mc_variance(s1, s2, spp, 1.0), mc_variance(r1, r2, spp, 1.0), f"C2 full size {flavor} {generator}")


def test_registered_host_bins_zero_copy_path(ctx):
    """vb200_host_register: the kernel applies '+=' / '=' to the caller's pinned bins in place; same bits as the staged path, shards included"""
    from viltrum_b200 import _capi, Vb200Error
    res, spp = [96, 80], 37
    nb = res[0] * res[1]
    rng = _rng(None, "shade4_16")
    init = np.linspace(-1, 1, nb).astype(np.float32)
    for flavor in (_capi.MC_PER_BIN, _capi.PER_BIN_MC):
        staged = init.copy()
        ctx.mc_per_bin("shade4_16", staged, res, rng, spp, 5, flavor)
        pinned = init.copy()
        ctx.host_register(pinned)
        try:
            ctx.mc_per_bin("shade4_16", pinned, res, rng, spp, 5, flavor, shard=(0, 3000))
            ctx.mc_per_bin("shade4_16", pinned, res, rng, spp, 5, flavor, shard=(3000, nb))
        finally:
            ctx.host_unregister(pinned)
        assert np.array_equal(staged.view(np.uint32), pinned.view(np.uint32))
    from viltrum_b200 import RangeInfinite
    a = np.zeros(64, np.float32); b = np.zeros(64, np.float32)
    ctx.mc_per_bin_inf("walk", a, [8, 8], RangeInfinite(), 32, 3)
    ctx.host_register(b)
    ctx.mc_per_bin_inf("walk", b, [8, 8], RangeInfinite(), 32, 3)
    with pytest.raises(Vb200Error):
        ctx.host_register(b)            # twice
    ctx.host_unregister(b)
    with pytest.raises(Vb200Error):
        ctx.host_unregister(b)          # not registered any more
    assert np.array_equal(a, b)
