// Counter-based Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel Random Numbers: As Easy as 1, 2, 3", SC'11).
// The reference has no counter-based RNG (its per-bin streams are std::mt19937 reseeded per bin,
// reference src/monte-carlo/monte-carlo-per-bin-parallel.h:50-58); a GPU wants a stateless generator keyed by
// (seed; bin, sample, draw block) so that every lane can jump straight to its own numbers and results do not
// depend on how the bin grid is sharded across CTAs or GPUs.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define VB200_HD __host__ __device__ __forceinline__
#else
#define VB200_HD inline
#endif

namespace viltrum { namespace b200 {

struct u32x4 { uint32_t x, y, z, w; };

VB200_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return uint32_t((uint64_t(a) * uint64_t(b)) >> 32);
#endif
}

template<int ROUNDS = 10>
VB200_HD u32x4 philox4x32(u32x4 c, uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        uint32_t hi0 = mulhi32(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = mulhi32(M1, c.z), lo1 = M1 * c.z;
        c = u32x4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
        k0 += W0; k1 += W1;
    }
    return c;
}

// The same function with the two products of the FIRST round handed in: p0 = M0 * c.x and p1 = M1 * c.z as 64-bit values.  A caller
// whose c.x is fixed over many calls keeps p0 in registers, and one whose c.z counts up by one gets the next p1 with a 64-bit add
// (philox_m1()) — two of the twenty 32x32->64 multiplies, the expensive instructions of this generator on the GPU (IMAD.WIDE), go away.
VB200_HD constexpr uint64_t philox_m0() { return 0xD2511F53ull; }
VB200_HD constexpr uint64_t philox_m1() { return 0xCD9E8D57ull; }
template<int ROUNDS = 10>
VB200_HD u32x4 philox4x32_from_products(uint64_t p0, uint64_t p1, uint32_t cy, uint32_t cw, uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    u32x4 c{uint32_t(p1 >> 32) ^ cy ^ k0, uint32_t(p1), uint32_t(p0 >> 32) ^ cw ^ k1, uint32_t(p0)};
    k0 += W0; k1 += W1;
#pragma unroll
    for (int r = 1; r < ROUNDS; ++r) {
        uint32_t hi0 = mulhi32(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = mulhi32(M1, c.z), lo1 = M1 * c.z;
        c = u32x4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
        k0 += W0; k1 += W1;
    }
    return c;
}

// Key schedule held by the caller (k0 + r*W0, k1 + r*W1): a kernel that calls the generator in a loop computes it once, behind an
// opaque move so that the compiler keeps the twenty words in registers instead of re-deriving them with 18 adds per call.
template<int ROUNDS = 10> struct PhiloxKeys { uint32_t k0[ROUNDS], k1[ROUNDS]; };
template<int ROUNDS = 10>
VB200_HD PhiloxKeys<ROUNDS> philox_key_schedule(uint32_t k0, uint32_t k1) {
    constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    PhiloxKeys<ROUNDS> ks;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        uint32_t a = k0 + uint32_t(r) * W0, b = k1 + uint32_t(r) * W1;
#if defined(__CUDA_ARCH__)
        asm volatile("mov.u32 %0, %0;" : "+r"(a));
        asm volatile("mov.u32 %0, %0;" : "+r"(b));
#endif
        ks.k0[r] = a; ks.k1[r] = b;
    }
    return ks;
}
template<int ROUNDS = 10>
VB200_HD u32x4 philox4x32_from_products(uint64_t p0, uint64_t p1, uint32_t cy, uint32_t cw, const PhiloxKeys<ROUNDS>& ks) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    u32x4 c{uint32_t(p1 >> 32) ^ cy ^ ks.k0[0], uint32_t(p1), uint32_t(p0 >> 32) ^ cw ^ ks.k1[0], uint32_t(p0)};
#pragma unroll
    for (int r = 1; r < ROUNDS; ++r) {
        uint32_t hi0 = mulhi32(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = mulhi32(M1, c.z), lo1 = M1 * c.z;
        c = u32x4{hi1 ^ c.y ^ ks.k0[r], lo1, hi0 ^ c.w ^ ks.k1[r], lo0};
    }
    return c;
}

// [0,1) from the top 24 bits — the mapping the reference's vendored generators use
// (reference src/rng/XoshiroCpp.hpp:650-654 FloatFromBits): never returns 1, exact in fp32.
VB200_HD float u01(uint32_t u) { return float(u >> 8) * 5.9604644775390625e-08f; }

VB200_HD float pick(const u32x4& r, int i) { return u01(i == 0 ? r.x : i == 1 ? r.y : i == 2 ? r.z : r.w); }

}} // namespace viltrum::b200
