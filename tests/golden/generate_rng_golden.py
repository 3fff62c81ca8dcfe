"""Generates tests/golden/rng_vectors.json: outputs of the reference's own vendored xoshiro128++ (src/rng/XoshiroCpp.hpp:531-589,
Xoshiro128PlusPlus) for a few states.  Run in the authoring container (needs /root/reference and g++):

    python tests/golden/generate_rng_golden.py
"""
import json
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
STATES = [[1, 2, 3, 4], [0x9E3779B9, 0, 0, 0], [0xFFFFFFFF] * 4, [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0, 0, 0, 1]]
SRC = r"""
#include "rng/XoshiroCpp.hpp"
#include <cstdio>
#include <cstdlib>
int main(int argc, char** argv) {
    XoshiroCpp::Xoshiro128PlusPlus::state_type s{ (uint32_t)strtoul(argv[1],0,0), (uint32_t)strtoul(argv[2],0,0), (uint32_t)strtoul(argv[3],0,0), (uint32_t)strtoul(argv[4],0,0) };
    XoshiroCpp::Xoshiro128PlusPlus g(s);
    for (int i = 0; i < 40; ++i) printf("%u\n", (unsigned)g());
}
"""
with tempfile.TemporaryDirectory() as d:
    open(os.path.join(d, "x.cpp"), "w").write(SRC)
    subprocess.run(["g++", "-std=c++17", "-O1", "-I/root/reference/src", "-o", os.path.join(d, "x"), os.path.join(d, "x.cpp")], check=True)
    out = []
    for st in STATES:
        r = subprocess.run([os.path.join(d, "x")] + [str(v) for v in st], capture_output=True, text=True, check=True)
        out.append({"state": st, "outputs": [int(x) for x in r.stdout.split()]})
json.dump({"generator": "xoshiro128++ 1.0 (reference src/rng/XoshiroCpp.hpp Xoshiro128PlusPlus)", "vectors": out}, open(os.path.join(HERE, "rng_vectors.json"), "w"))
print("wrote", len(out), "vectors")
