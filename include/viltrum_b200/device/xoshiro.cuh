// xoshiro128++ 1.0 (Blackman & Vigna, "Scrambled linear pseudorandom number generators", ACM TOMS 2021): 128 bits of state, period
// 2^128 - 1, all 32 output bits pass BigCrush; add / rotate / shift / xor only.  The reference vendors the same generator family
// (src/rng/XoshiroCpp.hpp:531-589, Xoshiro128PlusPlus) as one of its RNG template arguments.
//
// Here it is the per-bin STREAM generator of the fast sampler: one state per (bin, lane sub-stream), seeded with the four words of
// Philox4x32-10(key = seed, counter = (bin lo, bin hi, sub-stream, tag)) — the shape of the reference's per-bin seeding
// (monte-carlo-per-bin-parallel.h:50-58: a per-bin generator seeded from a master stream), with a counter-based master so that a
// bin's stream does not depend on how the grid is sharded.  Nine ALU-pipe instructions per 32-bit word (LOP3 folds the three-input
// xors) and none on the FMA pipe, against 20 IMAD.WIDE + 20 LOP3 per four words of Philox (see threefry.cuh for the pipe argument).
#pragma once
#include <stdint.h>
#include "philox.cuh"

namespace viltrum { namespace b200 {

struct Xoshiro128pp {
    uint32_t s0, s1, s2, s3;
    VB200_HD static uint32_t rotl(uint32_t x, int k) {
#if defined(__CUDA_ARCH__)
        return __funnelshift_l(x, x, k);
#else
        return (x << k) | (x >> (32 - k));
#endif
    }
    VB200_HD uint32_t next() {
        const uint32_t result = rotl(s0 + s3, 7) + s0;
        const uint32_t t = s1 << 9;
        // s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3; s2 ^= t; s3 = rotl(s3, 11)  — written so that every new word is one 3-input xor
        const uint32_t n1 = s1 ^ s2 ^ s0, n0 = s0 ^ s3 ^ s1, n2 = s2 ^ s0 ^ t, n3 = rotl(s3 ^ s1, 11);
        s0 = n0; s1 = n1; s2 = n2; s3 = n3;
        return result;
    }
    // state from a 128-bit seed block; the all-zero state (the generator's one fixed point) is mapped away
    VB200_HD void seed(const u32x4& k) {
        s0 = k.x; s1 = k.y; s2 = k.z; s3 = k.w;
        if ((s0 | s1 | s2 | s3) == 0u) s0 = 0x9E3779B9u;
    }
};

}} // namespace viltrum::b200
