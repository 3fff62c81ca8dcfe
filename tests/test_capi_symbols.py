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
