"""-m gpu: the C++ drop-in (include/viltrum_b200/viltrum.h) — the reference's own call sites recompiled with nvcc against
user-defined functors (examples/*.cu, built by __graft_entry__.build()).  Each program checks itself against the analytic
value the reference's examples print; the deterministic one is also compared bit for bit with the oracle."""
import os
import subprocess
import numpy as np
import pytest
from gpu_helpers import ctx   # noqa: F401
from helpers import assert_same_bits

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "examples", "bin")


def run(name, *args):
    exe = os.path.join(BIN, name)
    if not os.path.exists(exe):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "examples")], check=True)
    r = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, f"{name} failed:\n{r.stdout}\n{r.stderr}"
    return r.stdout


def test_readme_example_config1(ctx):
    out = run("montecarlo_2d")
    assert "Bin 9" in out and "single value" in out


def test_per_bin_shading_config2(ctx, port, tmp_path):
    f = tmp_path / "bins.f32"
    out = run("per_bin_shading", 128, 64, f)
    assert "mean of bins" in out
    got = np.fromfile(f, np.float32)
    ref, _, s1, s2 = port.mc_per_bin_parallel("shade4_64", [128, 128], [0] * 4, [1] * 4, 64, 3, record=True)
    from gpu_helpers import mc_variance
    var = mc_variance(s1, s2, 64, 1.0)
    z = (got.astype(np.float64) - ref) / np.sqrt(2 * var + 1e-30)
    assert np.mean(np.abs(z) > 3) < 0.01 and abs(np.mean(z)) < 0.05


def test_walk_config5(ctx):
    assert "walk: mean of bins" in run("walk", 128, 256)


def test_adaptive_newton_cotes_config3_bit_exact(ctx, port, tmp_path):
    f = tmp_path / "bins.f32"
    g = tmp_path / "tol.f32"
    out = run("adaptive_newton_cotes", 64, 20000, f, g)
    assert "20001 regions" in out
    got = np.fromfile(f, np.float32)
    want, _ = port.adaptive_iterations("smooth_edge2", "boole_simpson", "size_relative", 20000, [64, 64], [0, 0], [1, 1])
    assert_same_bits(got, want, "C++ drop-in (user functor, exact build) vs oracle")
    want_tol, _, _ = port.adaptive_tolerance("smooth_edge2", "boole_simpson", "default_absolute", 1e-7, [64, 64], [0, 0], [1, 1])
    assert_same_bits(np.fromfile(g, np.float32), want_tol, "integrator_adaptive_tolerance, C++ drop-in vs oracle")


def test_crespo2021_config4(ctx):
    assert "control variates: mean of bins" in run("crespo2021", 64, 2048, 16)


def test_fubini_family(ctx):
    out = run("fubini")
    assert "crespo2021_infinite<2>" in out and "fubini<1>(adaptive, monte_carlo(256)) finite" in out
