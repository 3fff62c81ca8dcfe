"""Builds libviltrum_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m viltrum_b200.build            # incremental
    python -m viltrum_b200.build --force

The library has no CPU path: loading it works anywhere, but vb200_create() fails without a CUDA device.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libviltrum_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"),
          "-Xptxas", "-v"]

# (source, extra flags).  builtin_exact.cu is the --fmad=false twin used by the bit-exact parity modes.
SOURCES = [
    ("capi.cu", []),
    ("comm.cu", []),
    ("regions.cu", ["--fmad=false"]),
    ("cv.cu", ["--fmad=false"]),
    ("refine_batched.cu", ["--fmad=false"]),
    ("builtin_fast.cu", []),
    ("builtin_fubini.cu", []),
    ("builtin_exact.cu", ["--fmad=false", "-DVILTRUM_B200_EXACT"]),
]


def _deps():
    deps = []
    for base in (os.path.join(ROOT, "include"), CSRC):
        for d, _, files in os.walk(base):
            deps += [os.path.join(d, f) for f in files if f.endswith((".h", ".cuh"))]
    return deps


def _compile(src, extra, force, newest_header):
    s = os.path.join(CSRC, src)
    o = os.path.join(OBJ, src.replace(".cu", ".o"))
    if not force and os.path.exists(o) and os.path.getmtime(o) > max(os.path.getmtime(s), newest_header):
        return o, ""
    cmd = [NVCC] + ARCH + COMMON + extra + ["-c", s, "-o", o]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    with open(o.replace(".o", ".ptxas.txt"), "w") as f:
        f.write(r.stderr)
    return o, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    newest = max(os.path.getmtime(p) for p in _deps())
    with ThreadPoolExecutor(max_workers=4) as ex:
        results = list(ex.map(lambda t: _compile(t[0], t[1], force, newest), SOURCES))
    objs = [o for o, _ in results]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        for _, log in results:
            if log:
                print(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
