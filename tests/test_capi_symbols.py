"""CPU: the C-ABI library loads, exports every symbol include/viltrum_b200.h declares, and refuses to run without a
GPU (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "viltrum_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(vb200_[a-z0-9_]+)\s*\(", text))
    names -= {"vb200_launch_fn"}
    return sorted(names)


def test_header_and_binding_agree():
    from viltrum_b200 import _capi
    assert declared_symbols() == sorted(_capi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from viltrum_b200 import _capi
    L = _capi.lib()
    for name in declared_symbols():
        assert hasattr(L, name), f"libviltrum_b200.so does not export {name}"


def test_builtin_integrands_registered():
    from viltrum_b200 import _capi, builtin_names
    names = builtin_names()
    for want in ("x2y2", "shade4_64", "shade4_16", "shade5_64", "smooth_edge2", "walk", "decay"):
        assert want in names
    L = _capi.lib()
    for n in names:
        assert L.vb200_builtin_integrand(n.encode(), 0) and L.vb200_builtin_integrand(n.encode(), 1)
    assert not L.vb200_builtin_integrand(b"nope", 0)


def test_xoshiro128pp_matches_the_reference_vendored_generator():
    """vb200_xoshiro128pp (the per-bin stream generator of the default sampler) against outputs of the reference's own
    Xoshiro128PlusPlus (src/rng/XoshiroCpp.hpp:531-589), tests/golden/rng_vectors.json (generator script beside it)"""
    import json
    from viltrum_b200 import _capi
    L = _capi.lib()
    vec = json.load(open(os.path.join(ROOT, "tests", "golden", "rng_vectors.json")))["vectors"]
    assert len(vec) >= 5
    for v in vec:
        st = (ctypes.c_uint32 * 4)(*v["state"]); n = len(v["outputs"]); out = (ctypes.c_uint32 * n)()
        L.vb200_xoshiro128pp(st, 17, out)                       # in two calls: the state carries over
        L.vb200_xoshiro128pp(st, n - 17, ctypes.cast(ctypes.addressof(out) + 17 * 4, ctypes.POINTER(ctypes.c_uint32)))
        assert list(out) == v["outputs"]


def test_threefry_known_answers():
    # Random123 kat_vectors for threefry4x32 (20 and 13 rounds); the 12-round variant of the experiment harness shares the code
    from viltrum_b200 import _capi
    L = _capi.lib()
    pi = [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]; pik = [0xa4093822, 0x299f31d0, 0x082efa98, 0xec4e6c89]
    kat = [(20, [0] * 4, [0] * 4, [0x9c6ca96a, 0xe17eae66, 0xfc10ecd4, 0x5256a7d8]),
           (20, [0xffffffff] * 4, [0xffffffff] * 4, [0x2a881696, 0x57012287, 0xf6c7446e, 0xa16a6732]),
           (20, pi, pik, [0x59cd1dbb, 0xb8879579, 0x86b5d00c, 0xac8b6d84]),
           (13, [0] * 4, [0] * 4, [0x531c7e4f, 0x39491ee5, 0x2c855a92, 0x3d6abf9a])]
    for rounds, c, k, want in kat:
        o = (ctypes.c_uint32 * 4)()
        assert L.vb200_threefry4x32(rounds, (ctypes.c_uint32 * 4)(*c), (ctypes.c_uint32 * 4)(*k), o) == 0
        assert list(o) == want
    assert L.vb200_threefry4x32(7, (ctypes.c_uint32 * 4)(), (ctypes.c_uint32 * 4)(), (ctypes.c_uint32 * 4)()) != 0


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    from viltrum_b200 import _capi
    L = _capi.lib()
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        c = (ctypes.c_uint32 * 4)(*ctr); k = (ctypes.c_uint32 * 2)(*key); o = (ctypes.c_uint32 * 4)()
        L.vb200_philox4x32_10(c, k, o)
        assert tuple(o) == want


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from viltrum_b200 import Context, Vb200Error
    with pytest.raises(Vb200Error) as e:
        Context(0)
    assert e.value.status == -1 and "no CPU fallback" in str(e.value)


def test_no_product_module_touches_the_oracle():
    # the oracle is test infrastructure: nothing under viltrum_b200/ or include/ may reference it
    bad = []
    for base in ("viltrum_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in d.split(os.sep):
                continue
            for f in files:
                if f.endswith((".py", ".h", ".cuh", ".cu")):
                    t = open(os.path.join(d, f), errors="ignore").read()
                    if re.search(r"pyoracle|liboracle|oracle_api|#include\s+\"[^\"]*oracle/", t):
                        bad.append(os.path.join(d, f))
    assert not bad, bad


def test_philox_from_cached_products_equals_philox(tmp_path):
    """philox4x32_from_products (first-round products kept in registers by the walk kernel: M0 * bin per bin, M1 * sample advanced
    by a 64-bit add) is the same function as philox4x32<10>: host build of the header, 10^6 random counters and sample runs."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "p.cpp"
    src.write_text(r'''
#include <cstdio>
#include <cstdint>
#include "viltrum_b200/device/philox.cuh"
using namespace viltrum::b200;
int main() {
    uint64_t x = 88172645463325252ull; auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return uint32_t(x >> 16); };
    for (int i = 0; i < 1000000; ++i) {
        const uint32_t b0 = rnd(), b1 = rnd(), blk = rnd() & 15u, k0 = rnd(), k1 = rnd();
        uint32_t s = (i & 1) ? rnd() : 0u;
        uint64_t p1 = uint64_t(s) * philox_m1();
        const uint64_t p0 = uint64_t(b0) * philox_m0();
        for (int j = 0; j < 3; ++j, ++s, p1 += philox_m1()) {
            if (s == 0xffffffffu) break;
            const u32x4 a = philox4x32<10>(u32x4{b0, b1, s, blk}, k0, k1);
            const u32x4 b = philox4x32_from_products<10>(p0, p1, b1, blk, k0, k1);
            if (a.x != b.x || a.y != b.y || a.z != b.z || a.w != b.w) { std::printf("mismatch at %d\n", i); return 1; }
        }
    }
    std::printf("ok\n"); return 0;
}
''')
    exe = tmp_path / "p"
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout
