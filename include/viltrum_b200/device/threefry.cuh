// Counter-based Threefry4x32-R (Salmon, Moraes, Dror, Shaw: "Parallel Random Numbers: As Easy as 1, 2, 3", SC'11; the
// Threefish block cipher of Skein with the tweak dropped and the rotation constants re-searched for 32-bit words).
// R = 20 is the paper's default, R = 12 the smallest variant the paper lists as Crush-resistant (Table 2, with safety margin;
// 4x32 fails below 9 rounds).  Known-answer vectors (Random123 kat_vectors, R = 20 and R = 13) are checked on the host in
// tests/test_rng_kat.py through vb200_threefry4x32.
//
// Why a second generator next to Philox: add / rotate / xor run on the ALU pipe of an SM sub-partition, which the per-bin sampler
// leaves two-thirds idle, while Philox's 32x32->64 multiplies (IMAD.WIDE) share the FMA-heavy pipe with the integrand's FFMA2 and
// do not overlap with them (profiles/pipes_r1.txt).  The per-bin sampler draws part of every sample group from each generator so
// that both pipes fill (device/mc_per_bin.cuh, GroupDraws).
#pragma once
#include <stdint.h>
#include "philox.cuh"

namespace viltrum { namespace b200 {

VB200_HD uint32_t rotl32(uint32_t x, int r) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(x, x, r);
#else
    return (x << r) | (x >> (32 - r));
#endif
}

// key schedule: ks[0..3] = key, ks[4] = 0x1BD11BDA ^ key[0] ^ key[1] ^ key[2] ^ key[3] (Skein's parity word)
struct ThreefryKeys { uint32_t ks[5]; };
VB200_HD ThreefryKeys threefry_key_schedule(uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3) {
    ThreefryKeys t;
    t.ks[0] = k0; t.ks[1] = k1; t.ks[2] = k2; t.ks[3] = k3; t.ks[4] = 0x1BD11BDAu ^ k0 ^ k1 ^ k2 ^ k3;
    return t;
}

template<int ROUNDS = 12>
VB200_HD u32x4 threefry4x32(u32x4 c, const ThreefryKeys& t) {
    // rotation constants R_32x4_{round % 8}_{0,1}
    constexpr int R0[8] = {10, 11, 13, 23, 6, 17, 25, 18};
    constexpr int R1[8] = {26, 21, 27, 5, 20, 11, 10, 20};
    uint32_t x0 = c.x + t.ks[0], x1 = c.y + t.ks[1], x2 = c.z + t.ks[2], x3 = c.w + t.ks[3];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        if ((r & 1) == 0) {
            x0 += x1; x1 = rotl32(x1, R0[r & 7]); x1 ^= x0;
            x2 += x3; x3 = rotl32(x3, R1[r & 7]); x3 ^= x2;
        } else {
            x0 += x3; x3 = rotl32(x3, R0[r & 7]); x3 ^= x0;
            x2 += x1; x1 = rotl32(x1, R1[r & 7]); x1 ^= x2;
        }
        if ((r & 3) == 3) {                       // key injection after every fourth round
            const int s = r / 4 + 1;
            x0 += t.ks[s % 5]; x1 += t.ks[(s + 1) % 5]; x2 += t.ks[(s + 2) % 5]; x3 += t.ks[(s + 3) % 5] + uint32_t(s);
        }
    }
    return u32x4{x0, x1, x2, x3};
}

}} // namespace viltrum::b200
