"""-m gpu: infinite-dimensional per-bin Monte Carlo (rows a5/a20) and the global scatter sampler (row a6)."""
import numpy as np
import pytest
from gpu_helpers import ctx, assert_statistically_equal, mc_variance   # noqa: F401
from helpers import load_golden, f32, assert_same_bits

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("integ,res,spp,rmin,rmax", [("walk", [64, 64], 256, (), ()), ("walk", [100], 64, (), ()),
                                                     ("decay", [16, 16], 512, (), ()), ("walk", [24, 20], 128, (0.1, 0.2, 0.0), (0.9, 0.7, 1.0)),
                                                     ("walk", [5, 4, 3], 200, (), ())])
def test_statistical_parity(ctx, port, integ, res, spp, rmin, rmax):
    from viltrum_b200 import RangeInfinite
    rng = RangeInfinite(list(rmin), list(rmax))
    nb = int(np.prod(res))
    g = np.zeros(nb, np.float32); s1 = np.zeros(nb, np.float32); s2 = np.zeros(nb, np.float32)
    ctx.mc_per_bin_inf(integ, g, res, rng, spp, 77, sum_f=s1, sum_f2=s2)
    vol = float(np.prod(np.asarray(rmax, np.float32) - np.asarray(rmin, np.float32))) if len(rmin) else 1.0
    if len(res) <= 2:
        r, r1, r2, _, _ = port.mc_per_bin_parallel_inf(integ, res, spp, 3, rmin, rmax, record=True)
    else:
        r = np.zeros(nb, np.float32); r1 = np.zeros(nb, np.float32); r2 = np.zeros(nb, np.float32)
        ctx.mc_per_bin_inf(integ, r, res, rng, spp, 78, sum_f=r1, sum_f2=r2)
    assert_statistically_equal(g, r, mc_variance(s1, s2, spp, vol), mc_variance(r1, r2, spp, vol), f"{integ} {res}")


def test_decay_analytic(ctx):
    """reference main/doc/montecarlo-infd.cc:32 — geometric series, decay/(1-decay) = 3"""
    from viltrum_b200 import integrate, monte_carlo_per_bin_parallel, range_primary_infinite
    bins = np.zeros(64, np.float32)
    integrate(monte_carlo_per_bin_parallel(8192, seed=2), bins, None, "decay", range_primary_infinite(), ctx=ctx)
    assert abs(float(bins.mean()) - 3.0) < 0.05


@pytest.mark.parametrize("integ,res,spp,rmin,rmax", [("walk", [12, 10], 16, (), ()), ("decay", [33], 8, (), ()),
                                                     ("walk", [6, 5], 24, (0.1, 0.2, 0.0), (0.9, 0.7, 1.0))])
def test_replay_bit_exact(ctx, port, integ, res, spp, rmin, rmax):
    from viltrum_b200 import RangeInfinite
    nb = int(np.prod(res))
    init = np.linspace(0.5, 2, nb).astype(np.float32)
    want, _, _, lens, elems = port.mc_per_bin_parallel_inf(integ, res, spp, 5, rmin, rmax, bins=init, record=True)
    offsets = np.concatenate([[0], np.cumsum(lens.astype(np.uint64))]).astype(np.uint64)
    got = init.copy()
    ctx.mc_per_bin_inf_replay(integ, got, res, RangeInfinite(list(rmin), list(rmax)), spp, offsets, np.ascontiguousarray(elems))
    assert_same_bits(got, want, f"{integ} replay")


def test_replay_golden_reference_vectors(ctx):
    from viltrum_b200 import RangeInfinite
    n = 0
    for v in load_golden():
        if v["path"] not in ("mc_per_bin_parallel_inf", "per_bin_parallel_mc_inf"):
            continue
        from viltrum_b200 import _capi
        lens = np.asarray(v["lens"], np.uint64)
        offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        got = np.zeros(int(np.prod(v["res"])), np.float32)
        ctx.mc_per_bin_inf_replay(v["integrand"], got, v["res"], RangeInfinite(v["rmin"], v["rmax"]), v["spp"], offsets,
                                  np.ascontiguousarray(f32(v["elems"])), flavor=_capi.MC_PER_BIN if v["path"] == "mc_per_bin_parallel_inf" else _capi.PER_BIN_MC)
        assert_same_bits(got, f32(v["bins"]), f"{v['integrand']} {v['res']}")
        n += 1
    assert n == 8


def test_replay_detects_short_sequences(ctx, port):
    from viltrum_b200 import RangeInfinite, Vb200Error
    _, _, _, lens, elems = port.mc_per_bin_parallel_inf("walk", [4], 4, 5, record=True)
    offsets = np.concatenate([[0], np.cumsum(lens.astype(np.uint64))]).astype(np.uint64)
    offsets[1:] -= 1       # every path one element short
    with pytest.raises(Vb200Error):
        ctx.mc_per_bin_inf_replay("walk", np.zeros(4, np.float32), [4], RangeInfinite(), 4, offsets, np.ascontiguousarray(elems))


def test_walk_sharding_and_full_size_mean(ctx):
    """config 5 shape at 1/16 size (512x512 bins, 256 spp): mean = 1.0133, shards reproduce the whole"""
    import torch
    from viltrum_b200 import RangeInfinite
    res, spp = [512, 512], 256
    d = torch.zeros(512 * 512, dtype=torch.float32, device="cuda")
    ctx.mc_per_bin_inf("walk", d, res, RangeInfinite(), spp, 0); ctx.synchronize()
    a = d.cpu().numpy()
    assert abs(float(a.mean(dtype=np.float64)) - 1.0133) < 2e-3
    p = torch.zeros(512 * 512, dtype=torch.float32, device="cuda")
    for lo, hi in ((0, 100000), (100000, 100001), (100001, 512 * 512)):
        ctx.mc_per_bin_inf("walk", p, res, RangeInfinite(), spp, 0, shard=(lo, hi))
    ctx.synchronize()
    assert_same_bits(p.cpu().numpy(), a, "sharded walk")


def test_global_scatter_readme_example(ctx, port):
    """BASELINE.json config 1 — README example: monte_carlo(8192), f = x^2+y^2 over [0,1]^2 into 10 bins (std::vector<float>)."""
    from viltrum_b200 import integrate, monte_carlo, Range
    bins = np.zeros(10, np.float32)
    integrate(monte_carlo(8192, seed=0), bins, None, "x2y2", Range([0, 0], [1, 1]), ctx=ctx)
    analytic = np.array([(3 * k * k + 3 * k + 1) / 300 + 1 / 3 for k in range(10)])
    ref = port.monte_carlo("x2y2", [10], [0, 0], [1, 1], 8192, 0)
    # per-bin sigma of this estimator ~ 0.03; both the reference and the GPU sit around the analytic value
    assert np.max(np.abs(bins - analytic)) < 0.15 and np.max(np.abs(ref - analytic)) < 0.15
    big = np.zeros(10, np.float32)
    integrate(monte_carlo(1 << 24, seed=1), big, None, "x2y2", Range([0, 0], [1, 1]), ctx=ctx)
    assert np.allclose(big, analytic, atol=4e-3)
    acc = np.full(10, 5.0, np.float32)
    integrate(monte_carlo(1 << 24, seed=1), acc, None, "x2y2", Range([0, 0], [1, 1]), ctx=ctx)
    assert np.allclose(acc - 5.0, big, atol=1e-5)                     # '+=' (monte-carlo.h:59)


def test_global_scatter_split_samples_sum_to_whole(ctx):
    """split-bin mode: sample ranges drawn by different calls/GPUs add up to the single-call estimate (the allreduce path)"""
    import torch
    from viltrum_b200 import Range
    res, n = [64, 64], 1 << 22
    rng = Range([0] * 4, [1] * 4)
    whole = torch.zeros(4096, dtype=torch.float32, device="cuda")
    ctx.monte_carlo("shade4_16", whole, res, rng, n, 9)
    parts = [torch.zeros(4096, dtype=torch.float32, device="cuda") for _ in range(4)]
    for i, p in enumerate(parts):
        ctx.monte_carlo("shade4_16", p, res, rng, n, 9, shard=(i * n // 4, (i + 1) * n // 4))
    ctx.synchronize()
    tot = sum(p.double() for p in parts)
    assert torch.allclose(tot, whole.double(), rtol=1e-4, atol=1e-5)      # float atomics: order differs, values agree
    assert abs(float(whole.mean()) - 0.1430) < 5e-3


def test_wavefront_kernel_equals_generic_kernel(ctx):
    """'walk' carries the state-machine form AND its element counts (block-fed wavefront kernel: one Philox block per lane and iteration,
    immediate refill), 'walk_steps' the state machine only (wavefront kernel with batched refill through the general iterator),
    'walk_plain' only operator()(seq) (generic per-lane kernel): same Philox elements, same per-lane summation order ->
    bit-identical bins, also over explicit range entries (in block 0: block kernel; beyond: the launcher falls back) and both flavors."""
    from viltrum_b200 import RangeInfinite, _capi
    cases = [([64, 48], 256, (), ()), ([100], 37, (), ()), ([9, 7, 5], 64, (), ()), ([20, 16], 128, (0.1, 0.2, 0.0), (0.9, 0.7, 1.0)),
             ([12, 10], 33, (0.1, 0.2, 0.0, 0.0, 0.25), (0.9, 0.7, 1.0, 1.0, 0.75)), ([5], 1000, (-1.0,), (2.0,))]
    for res, spp, rmin, rmax in cases:
        nb = int(np.prod(res))
        rng = RangeInfinite(list(rmin), list(rmax))
        for flavor in (_capi.MC_PER_BIN, _capi.PER_BIN_MC):
            out = {}
            for name in ("walk", "walk_steps", "walk_plain"):
                out[name] = np.zeros(nb, np.float32)
                ctx.mc_per_bin_inf(name, out[name], res, rng, spp, 5, flavor=flavor)
            assert_same_bits(out["walk"], out["walk_plain"], f"block-fed vs generic {res} {rmin}")
            assert_same_bits(out["walk_steps"], out["walk_plain"], f"wavefront vs generic {res} {rmin}")
            d = np.zeros(nb, np.float32); dp = np.zeros(nb, np.float32)       # 'decay': begin() reads nothing, two elements per round
            ctx.mc_per_bin_inf("decay", d, res, rng, spp, 9, flavor=flavor)
            ctx.mc_per_bin_inf("decay_plain", dp, res, rng, spp, 9, flavor=flavor)
            assert_same_bits(d, dp, f"block-fed vs generic decay {res} {rmin}")


@pytest.mark.parametrize("integ,res,spp,rmin,rmax", [("walk", [48, 40], 256, (), ()), ("decay", [64], 512, (), ()),
                                                     ("walk", [20, 16], 128, (0.1, 0.2, 0.0), (0.9, 0.7, 1.0))])
def test_wrapper_spelling_over_infinite_range_a20(ctx, port, integ, res, spp, rmin, rmax):
    """integrator_per_bin_parallel(monte_carlo(spp,seed)) over RangeInfinite ('=', scaled by the bin box's
    own volume) — statistical parity with the oracle and bit-exact replay of the oracle's recorded sequences."""
    from viltrum_b200 import RangeInfinite, integrate, integrator_per_bin_parallel, monte_carlo, _capi
    rng = RangeInfinite(list(rmin), list(rmax))
    nb = int(np.prod(res))
    g = np.full(nb, 9.0, np.float32); s1 = np.zeros(nb, np.float32); s2 = np.zeros(nb, np.float32)
    integrate(integrator_per_bin_parallel(monte_carlo(spp, seed=21)), g, res, integ, rng, ctx=ctx, sum_f=s1, sum_f2=s2)
    r, r1, r2, lens, elems = port.per_bin_parallel_mc_inf(integ, res, spp, 4, rmin, rmax, record=True)
    vol = float(np.prod(np.asarray(rmax, np.float32) - np.asarray(rmin, np.float32))) if len(rmin) else 1.0
    assert_statistically_equal(g, r, mc_variance(s1, s2, spp, vol), mc_variance(r1, r2, spp, vol), f"a20 {integ} {res}")
    offsets = np.concatenate([[0], np.cumsum(lens.astype(np.uint64))]).astype(np.uint64)
    got = np.full(nb, 3.0, np.float32)
    ctx.mc_per_bin_inf_replay(integ, got, res, rng, spp, offsets, np.ascontiguousarray(elems), flavor=_capi.PER_BIN_MC)
    assert_same_bits(got, r, "a20 replay")


@pytest.mark.parametrize("integ,res,rmin,rmax", [("walk", [16, 12], (), ()), ("decay", [10], (), ()), ("walk", [8, 8], (0.1, 0.2, 0.0), (0.9, 0.7, 1.0))])
def test_global_scatter_over_infinite_range(ctx, port, integ, res, rmin, rmax):
    """monte_carlo(n,seed) over RangeInfinite (monte-carlo.h:65-84): bin from the first dimbins sequence elements"""
    from viltrum_b200 import RangeInfinite, integrate, monte_carlo
    rng = RangeInfinite(list(rmin), list(rmax))
    nb = int(np.prod(res)); n = 1 << 22
    g = np.zeros(nb, np.float32)
    integrate(monte_carlo(n, seed=5), g, res, integ, rng, ctx=ctx)
    ref_n = 400000
    r = port.monte_carlo_inf(integ, res, ref_n, 3, rmin, rmax)
    # the oracle run is the noisy one: per-bin sigma ~ |f| * sqrt(nb/ref_n)
    sigma = (np.abs(r).mean() + 1.0) * np.sqrt(nb / ref_n) * 1.5
    assert np.max(np.abs(g - r)) < 6 * sigma, (np.max(np.abs(g - r)), sigma)
    assert abs(float(g.mean()) - float(r.mean())) < 6 * sigma / np.sqrt(nb) + 1e-3
    acc = np.full(nb, 2.0, np.float32)
    integrate(monte_carlo(n, seed=5), acc, res, integ, rng, ctx=ctx)
    assert np.allclose(acc - 2.0, g, rtol=1e-4, atol=1e-4)          # '+=' ; float atomics reorder sums


def test_window_kernel_equals_tile_kernel(ctx, monkeypatch):
    """One lane per bin (what every large grid gets; forced here with VB200_LANES_PER_BIN=1): the two-tile window kernel
    (walk_block_window_kernel: a lane moves on to its bin of the warp's next tile instead of idling) against the tile-at-a-time
    block kernel and the generic kernel — same elements, same per-bin summation order -> bit-identical bins and moments; device bins,
    host bins (end-to-end path: chunk flags raised per retired tile), ragged shards, both flavors, accumulate, more warps than tiles."""
    import torch
    from viltrum_b200 import RangeInfinite, _capi
    monkeypatch.setenv("VB200_LANES_PER_BIN", "1")
    cases = [([64, 48], 64, (), ()), ([1000], 37, (), ()), ([9, 7, 5], 40, (), ()), ([50, 41], 96, (0.1, 0.2, 0.0), (0.9, 0.7, 1.0)),
             ([3], 500, (-1.0,), (2.0,)), ([256, 200], 32, (), ())]
    for res, spp, rmin, rmax in cases:
        nb = int(np.prod(res))
        rng = RangeInfinite(list(rmin), list(rmax))
        shards = [None, (0, nb // 3 + 1), (nb // 3 + 1, nb)]
        for flavor in (_capi.MC_PER_BIN, _capi.PER_BIN_MC):
            for name, plain in (("walk", "walk_plain"), ("decay", "decay_plain")):
                got = {}
                for window in ("1", "0"):
                    monkeypatch.setenv("VB200_WALK_WINDOW", window)
                    h = np.full(nb, 0.25, np.float32); s1 = np.zeros(nb, np.float32); s2 = np.zeros(nb, np.float32)
                    ctx.mc_per_bin_inf(name, h, res, rng, spp, 11, flavor=flavor, sum_f=s1, sum_f2=s2)        # host bins
                    d = torch.full((nb,), 0.25, dtype=torch.float32, device="cuda")
                    for sh in shards[1:]:
                        ctx.mc_per_bin_inf(name, d, res, rng, spp, 11, flavor=flavor, shard=sh)                # device bins, ragged shards
                    ctx.synchronize()
                    got[window] = (h, s1, s2, d.cpu().numpy())
                ref = np.full(nb, 0.25, np.float32)
                ctx.mc_per_bin_inf(plain, ref, res, rng, spp, 11, flavor=flavor)
                for k, what in enumerate(("host bins", "sum f", "sum f^2", "sharded device bins")):
                    assert_same_bits(got["1"][k], got["0"][k], f"window vs tile kernel: {what} {name} {res} {rmin} flavor {flavor}")
                assert_same_bits(got["1"][0], ref, f"window vs generic kernel {name} {res} {rmin} flavor {flavor}")
                assert_same_bits(got["1"][3], ref, f"window (sharded) vs generic kernel {name} {res} {rmin} flavor {flavor}")


def test_full_size_per_bin_parity_against_the_reference(ctx):
    """BASELINE config 5 at FULL size (2048x2048 bins, 256 spp): every bin within 3 sigma of the unmodified reference's estimate of the same
    bin, each side with its own per-bin moments (oracle/_ref multi-threaded build; ~20 s of host time)."""
    import os
    import pyoracle
    from viltrum_b200 import RangeInfinite
    if not pyoracle.available("reference-mt"):
        pytest.skip("oracle/_ref/libviltrum_ref_mt.so was not built (needs /root/reference at build time)")
    O = pyoracle.load("reference-mt")
    O.set_threads(len(os.sched_getaffinity(0)))
    res, spp, nb = [2048, 2048], 256, 1 << 22
    ref, r1, r2 = O.mc_per_bin_parallel_inf("walk", res, spp, 77, record="moments")
    g = np.zeros(nb, np.float32); s1 = np.zeros(nb, np.float32); s2 = np.zeros(nb, np.float32)
    ctx.mc_per_bin_inf("walk", g, res, RangeInfinite(), spp, 5, sum_f=s1, sum_f2=s2)
    assert_statistically_equal(g, ref, mc_variance(s1, s2, spp, 1.0), mc_variance(r1, r2, spp, 1.0), "C5 full size")
