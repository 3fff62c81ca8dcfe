// Experiment harness (not product): packed FP32x2 (FFMA2) evaluation of the per-bin MC kernel on the C2 workload,
// plus pipe microbenchmarks (FFMA, FFMA2, IMAD.WIDE, mixes).  Build: see profiles/exp/Makefile-less recipe in run_packed.sh
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <array>
#include <cuda_runtime.h>
#include <stdint.h>

struct u32x4 { uint32_t x, y, z, w; };
template<int ROUNDS> __device__ __forceinline__ u32x4 philox(u32x4 c, uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = u32x4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
        k0 += W0; k1 += W1;
    }
    return c;
}

// ---- packed pair of floats -------------------------------------------------------------------------------------------
struct f2 { unsigned long long v; };
__device__ __forceinline__ f2 mk(float a, float b) { f2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f2 bc(float a) { return mk(a, a); }
__device__ __forceinline__ float lo(f2 a) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); return x; }
__device__ __forceinline__ float hi(f2 a) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); return y; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }

template<int K> __device__ __forceinline__ float shade4(const float (&x)[4]) {
    const float a = x[0]-.5f, b = x[1]-.5f;
    const float edge = .55f+.35f*(a*a-b*b)+.2f*a*b;
    const float vis = (x[2]+.5f*x[3]<edge)?1.0f:0.0f;
    const float t = x[2]*(1.0f-x[3]);
    float lobe = 1.0f/float(K);
#pragma unroll
    for (int k=K-2;k>=0;--k) lobe = lobe*t+1.0f/float(k+1);
    const float alb = .25f+.75f*x[0]*x[1];
    return vis*lobe*alb;
}
template<int K> __device__ __forceinline__ f2 shade4_2(const f2 (&x)[4]) {
    const f2 a = add2(x[0], bc(-.5f)), b = add2(x[1], bc(-.5f));
    const f2 d = sub2(mul2(a, a), mul2(b, b));
    const f2 edge = fma2(bc(.2f), mul2(a, b), fma2(bc(.35f), d, bc(.55f)));
    const f2 lhs = fma2(bc(.5f), x[3], x[2]);
    const f2 t = mul2(x[2], sub2(bc(1.0f), x[3]));
    f2 lobe = bc(1.0f/float(K));
#pragma unroll
    for (int k=K-2;k>=0;--k) lobe = fma2(lobe, t, bc(1.0f/float(k+1)));
    const f2 alb = fma2(mul2(bc(.75f), x[0]), x[1], bc(.25f));
    const f2 r = mul2(lobe, alb);
    return mk(lo(lhs) < lo(edge) ? lo(r) : 0.0f, hi(lhs) < hi(edge) ? hi(r) : 0.0f);
}

struct Args { uint32_t res0, res1; uint32_t spp; uint32_t k0, k1; float* out; uint64_t nbins; unsigned* counter; };

__device__ __forceinline__ void box(const Args& a, uint32_t bin, float (&lo_)[4], float (&ext)[4]) {
    const uint32_t p0 = bin % a.res0, p1 = bin / a.res0;
    const float d0 = 1.0f / float(a.res0), d1 = 1.0f / float(a.res1);
    lo_[0] = float(p0) * d0; ext[0] = float(p0 + 1) * d0 - lo_[0];
    lo_[1] = float(p1) * d1; ext[1] = float(p1 + 1) * d1 - lo_[1];
    lo_[2] = 0; ext[2] = 1; lo_[3] = 0; ext[3] = 1;
}
// CONV 0: (w>>8) I2F * ext24 + lo ; CONV 1: ((w & 0x7fffff)|0x3f800000) as float in [1,2): f*ext + (lo-ext)
template<int CONV> __device__ __forceinline__ float tofl(uint32_t w) {
    if (CONV == 0) return float(w >> 8);
    return __uint_as_float((w & 0x007fffffu) | 0x3f800000u);
}

// PACK 0: scalar, two samples in flight; PACK 1: f32x2 pair; PAIRS = pairs in flight (PACK 1)
template<int ROUNDS, int CONV, int PACK, int PAIRS, int LPB, int MINB>
__global__ void __launch_bounds__(256, MINB) k_mc(const Args a) {
    constexpr int G = 32 / LPB;
    const uint32_t lane = threadIdx.x & 31, sub = lane % LPB, grp = lane / LPB;
    const uint32_t ntiles = uint32_t(a.nbins / G);
    uint32_t tile;
    if (lane == 0) tile = atomicAdd(a.counter, 1u); tile = __shfl_sync(0xffffffffu, tile, 0);
    while (tile < ntiles) {
        const uint32_t bin = tile * G + grp;
        float lo_[4], ext[4]; box(a, bin, lo_, ext);
#pragma unroll
        for (int i = 0; i < 4; ++i) { if (CONV == 0) ext[i] *= 5.9604644775390625e-08f; else lo_[i] -= ext[i]; }
        float sum = 0.f;
        if (PACK == 0) {
            float sumb = 0.f;
            for (uint32_t s = sub; s < a.spp; s += 2 * LPB) {
                const u32x4 r0 = philox<ROUNDS>(u32x4{bin, 0u, s, 0u}, a.k0, a.k1);
                const u32x4 r1 = philox<ROUNDS>(u32x4{bin, 0u, s + LPB, 0u}, a.k0, a.k1);
                float x[4] = {fmaf(tofl<CONV>(r0.x), ext[0], lo_[0]), fmaf(tofl<CONV>(r0.y), ext[1], lo_[1]), fmaf(tofl<CONV>(r0.z), ext[2], lo_[2]), fmaf(tofl<CONV>(r0.w), ext[3], lo_[3])};
                float y[4] = {fmaf(tofl<CONV>(r1.x), ext[0], lo_[0]), fmaf(tofl<CONV>(r1.y), ext[1], lo_[1]), fmaf(tofl<CONV>(r1.z), ext[2], lo_[2]), fmaf(tofl<CONV>(r1.w), ext[3], lo_[3])};
                sum += shade4<64>(x); sumb += shade4<64>(y);
            }
            sum += sumb;
        } else {
            f2 acc[PAIRS];
#pragma unroll
            for (int p = 0; p < PAIRS; ++p) acc[p] = bc(0.f);
            for (uint32_t s = sub; s < a.spp; s += 2 * PAIRS * LPB) {
#pragma unroll
                for (int p = 0; p < PAIRS; ++p) {
                    const u32x4 r0 = philox<ROUNDS>(u32x4{bin, 0u, s + (2 * p) * LPB, 0u}, a.k0, a.k1);
                    const u32x4 r1 = philox<ROUNDS>(u32x4{bin, 0u, s + (2 * p + 1) * LPB, 0u}, a.k0, a.k1);
                    f2 x[4] = {fma2(mk(tofl<CONV>(r0.x), tofl<CONV>(r1.x)), bc(ext[0]), bc(lo_[0])),
                               fma2(mk(tofl<CONV>(r0.y), tofl<CONV>(r1.y)), bc(ext[1]), bc(lo_[1])),
                               fma2(mk(tofl<CONV>(r0.z), tofl<CONV>(r1.z)), bc(ext[2]), bc(lo_[2])),
                               fma2(mk(tofl<CONV>(r0.w), tofl<CONV>(r1.w)), bc(ext[3]), bc(lo_[3]))};
                    acc[p] = add2(acc[p], shade4_2<64>(x));
                }
            }
#pragma unroll
            for (int p = 0; p < PAIRS; ++p) sum += lo(acc[p]) + hi(acc[p]);
        }
#pragma unroll
        for (int off = LPB / 2; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if (sub == 0) a.out[bin] = sum * (1.0f / float(a.spp));
        if (lane == 0) tile = atomicAdd(a.counter, 1u); tile = __shfl_sync(0xffffffffu, tile, 0);
    }
}


// 4 samples per 3 Philox calls: samples 0..2 take the top 24 bits of every word of call 0..2, sample 3 is assembled from the three
// low bytes (PRMT), so no generated bit is thrown away (24 bits x 4 coordinates x 4 samples = 384 bits = 3 x 128)
template<int ROUNDS, int LPB, int MINB>
__global__ void __launch_bounds__(256, MINB) k_mc34(const Args a) {
    constexpr int G = 32 / LPB;
    const uint32_t lane = threadIdx.x & 31, sub = lane % LPB, grp = lane / LPB;
    const uint32_t ntiles = uint32_t(a.nbins / G);
    uint32_t tile;
    if (lane == 0) tile = atomicAdd(a.counter, 1u); tile = __shfl_sync(0xffffffffu, tile, 0);
    while (tile < ntiles) {
        const uint32_t bin = tile * G + grp;
        float lo_[4], ext[4]; box(a, bin, lo_, ext);
#pragma unroll
        for (int i = 0; i < 4; ++i) ext[i] *= 5.9604644775390625e-08f;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        for (uint32_t g = sub; g < a.spp / 4; g += LPB) {
            const u32x4 r0 = philox<ROUNDS>(u32x4{bin, 0u, g, 0u}, a.k0, a.k1);
            const u32x4 r1 = philox<ROUNDS>(u32x4{bin, 0u, g, 1u}, a.k0, a.k1);
            const u32x4 r2 = philox<ROUNDS>(u32x4{bin, 0u, g, 2u}, a.k0, a.k1);
            float x[4] = {fmaf(float(r0.x >> 8), ext[0], lo_[0]), fmaf(float(r0.y >> 8), ext[1], lo_[1]), fmaf(float(r0.z >> 8), ext[2], lo_[2]), fmaf(float(r0.w >> 8), ext[3], lo_[3])};
            float y[4] = {fmaf(float(r1.x >> 8), ext[0], lo_[0]), fmaf(float(r1.y >> 8), ext[1], lo_[1]), fmaf(float(r1.z >> 8), ext[2], lo_[2]), fmaf(float(r1.w >> 8), ext[3], lo_[3])};
            float z[4] = {fmaf(float(r2.x >> 8), ext[0], lo_[0]), fmaf(float(r2.y >> 8), ext[1], lo_[1]), fmaf(float(r2.z >> 8), ext[2], lo_[2]), fmaf(float(r2.w >> 8), ext[3], lo_[3])};
            // low bytes: (r0.b0 << 16) | (r1.b0 << 8) | r2.b0
            auto low = [] (uint32_t p, uint32_t q, uint32_t r) { return __byte_perm(__byte_perm(r, q, 0x7740), p, 0x7410) & 0x00ffffffu; };
            float w[4] = {fmaf(float(low(r0.x, r1.x, r2.x)), ext[0], lo_[0]), fmaf(float(low(r0.y, r1.y, r2.y)), ext[1], lo_[1]),
                          fmaf(float(low(r0.z, r1.z, r2.z)), ext[2], lo_[2]), fmaf(float(low(r0.w, r1.w, r2.w)), ext[3], lo_[3])};
            s0 += shade4<64>(x); s1 += shade4<64>(y); s2 += shade4<64>(z); s3 += shade4<64>(w);
        }
        float sum = (s0 + s1) + (s2 + s3);
#pragma unroll
        for (int off = LPB / 2; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if (sub == 0) a.out[bin] = sum * (1.0f / float(a.spp));
        if (lane == 0) tile = atomicAdd(a.counter, 1u); tile = __shfl_sync(0xffffffffu, tile, 0);
    }
}


// the same draw scheme feeding two packed pairs (FFMA2): samples (0,1) and (2,3)
template<int ROUNDS, int LPB, int MINB>
__global__ void __launch_bounds__(256, MINB) k_mc34p(const Args a) {
    constexpr int G = 32 / LPB;
    const uint32_t lane = threadIdx.x & 31, sub = lane % LPB, grp = lane / LPB;
    const uint32_t ntiles = uint32_t(a.nbins / G);
    uint32_t tile;
    if (lane == 0) tile = atomicAdd(a.counter, 1u); tile = __shfl_sync(0xffffffffu, tile, 0);
    while (tile < ntiles) {
        const uint32_t bin = tile * G + grp;
        float lo_[4], ext[4]; box(a, bin, lo_, ext);
#pragma unroll
        for (int i = 0; i < 4; ++i) ext[i] *= 5.9604644775390625e-08f;
        f2 acc0 = bc(0.f), acc1 = bc(0.f);
        for (uint32_t g = sub; g < a.spp / 4; g += LPB) {
            const u32x4 r0 = philox<ROUNDS>(u32x4{bin, 0u, g, 0u}, a.k0, a.k1);
            const u32x4 r1 = philox<ROUNDS>(u32x4{bin, 0u, g, 1u}, a.k0, a.k1);
            const u32x4 r2 = philox<ROUNDS>(u32x4{bin, 0u, g, 2u}, a.k0, a.k1);
            auto low = [] (uint32_t p, uint32_t q, uint32_t r) { return __byte_perm(__byte_perm(r, q, 0x7740), p, 0x7410) & 0x00ffffffu; };
            f2 x[4] = {fma2(mk(float(r0.x >> 8), float(r1.x >> 8)), bc(ext[0]), bc(lo_[0])), fma2(mk(float(r0.y >> 8), float(r1.y >> 8)), bc(ext[1]), bc(lo_[1])),
                       fma2(mk(float(r0.z >> 8), float(r1.z >> 8)), bc(ext[2]), bc(lo_[2])), fma2(mk(float(r0.w >> 8), float(r1.w >> 8)), bc(ext[3]), bc(lo_[3]))};
            f2 y[4] = {fma2(mk(float(r2.x >> 8), float(low(r0.x, r1.x, r2.x))), bc(ext[0]), bc(lo_[0])), fma2(mk(float(r2.y >> 8), float(low(r0.y, r1.y, r2.y))), bc(ext[1]), bc(lo_[1])),
                       fma2(mk(float(r2.z >> 8), float(low(r0.z, r1.z, r2.z))), bc(ext[2]), bc(lo_[2])), fma2(mk(float(r2.w >> 8), float(low(r0.w, r1.w, r2.w))), bc(ext[3]), bc(lo_[3]))};
            acc0 = add2(acc0, shade4_2<64>(x)); acc1 = add2(acc1, shade4_2<64>(y));
        }
        float sum = (lo(acc0) + hi(acc0)) + (lo(acc1) + hi(acc1));
#pragma unroll
        for (int off = LPB / 2; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if (sub == 0) a.out[bin] = sum * (1.0f / float(a.spp));
        if (lane == 0) tile = atomicAdd(a.counter, 1u); tile = __shfl_sync(0xffffffffu, tile, 0);
    }
}

// ---- pipe microbenchmarks ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ffma(float* out, int iters, float a) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = float(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], a, 0.0123f);
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += v[i];
    if (s == 12345.678f) out[0] = s;
}
__global__ void __launch_bounds__(256) k_ffma2(float* out, int iters, float a) {
    f2 v[8]; const f2 aa = bc(a);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = mk(float(threadIdx.x + i), float(i));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fma2(v[i], aa, bc(0.0123f));
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += lo(v[i]) + hi(v[i]);
    if (s == 12345.678f) out[0] = s;
}
// NW IMAD.WIDE + NF FFMA (NF2 FFMA2) per inner iteration, independent chains
template<int NW, int NF, int NF2>
__global__ void __launch_bounds__(256) k_mix(float* out, int iters, float a) {
    uint32_t w[NW > 0 ? NW : 1]; float v[NF > 0 ? NF : 1]; f2 p[NF2 > 0 ? NF2 : 1];
    for (int i = 0; i < NW; ++i) w[i] = threadIdx.x * 7 + i;
    for (int i = 0; i < NF; ++i) v[i] = float(threadIdx.x + i);
    for (int i = 0; i < NF2; ++i) p[i] = mk(float(threadIdx.x + i), float(i));
    const f2 aa = bc(a);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NW; ++i) { const unsigned long long m = (unsigned long long)w[i] * 0xD2511F53u; w[i] = uint32_t(m >> 32) ^ uint32_t(m); }
#pragma unroll
        for (int i = 0; i < NF; ++i) v[i] = fmaf(v[i], a, 0.0123f);
#pragma unroll
        for (int i = 0; i < NF2; ++i) p[i] = fma2(p[i], aa, bc(0.0123f));
    }
    float s = 0; for (int i = 0; i < NF; ++i) s += v[i]; for (int i = 0; i < NF2; ++i) s += lo(p[i]) + hi(p[i]);
    uint32_t x = 0; for (int i = 0; i < NW; ++i) x ^= w[i];
    if (s == 12345.678f || x == 0x12345u) out[0] = s;
}

template<class K> int occ_grid(K k, int sms) { int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 256, 0); return occ * sms; }
template<class L> float time_ms(L launch, int reps = 10) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch();
    cudaDeviceSynchronize();
    float tot = 0;
    for (int i = 0; i < reps; ++i) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); tot += ms; }
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
    return tot / reps;
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    Args a; a.res0 = 1024; a.res1 = 1024; a.spp = 64; a.k0 = 1; a.k1 = 2; a.nbins = 1u << 20;
    cudaMalloc(&a.out, a.nbins * 4); cudaMalloc(&a.counter, 4);
    const double evals = double(a.nbins) * a.spp;
    std::vector<float> h(a.nbins);
    auto report = [&](const char* name, float ms, int grid) {
        cudaMemcpy(h.data(), a.out, a.nbins * 4, cudaMemcpyDeviceToHost);
        double m = 0; for (float v : h) m += v; m /= a.nbins;
        printf("%-44s grid %5d  %8.3f ms  %7.1f Gevals/s  %5.1f TFLOP/s(155)  mean %.5f\n", name, grid, ms, evals / ms * 1e-6, evals * 155 / ms * 1e-9, m);
    };
#define RUN(NAME, ...) { auto k = k_mc<__VA_ARGS__>; int g = occ_grid(k, sms); \
        float ms = time_ms([&] { cudaMemsetAsync(a.counter, 0, 4); k<<<g, 256>>>(a); }); report(NAME, ms, g); }
    //            ROUNDS CONV PACK PAIRS LPB MINB
    RUN("scalar ILP2 i2f LPB1",            10, 0, 0, 1, 1, 1)
    RUN("scalar ILP2 lop LPB1",            10, 1, 0, 1, 1, 1)
    RUN("packed 1 pair i2f LPB1",          10, 0, 1, 1, 1, 1)
    RUN("packed 1 pair lop LPB1",          10, 1, 1, 1, 1, 1)
    RUN("packed 2 pairs lop LPB1",         10, 1, 1, 2, 1, 1)
    RUN("packed 1 pair lop LPB2",          10, 1, 1, 1, 2, 1)
    RUN("packed 1 pair lop LPB4",          10, 1, 1, 1, 4, 1)
    RUN("packed 2 pairs lop LPB4",         10, 1, 1, 2, 4, 1)
    RUN("packed 1 pair lop LPB1 minb4",    10, 1, 1, 1, 1, 4)
    RUN("packed 1 pair lop LPB1 minb6",    10, 1, 1, 1, 1, 6)
    RUN("packed 1 pair lop LPB1 minb8",    10, 1, 1, 1, 1, 8)
#define RUN34(NAME, ...) { auto k = k_mc34<__VA_ARGS__>; int g = occ_grid(k, sms); \
        float ms = time_ms([&] { cudaMemsetAsync(a.counter, 0, 4); k<<<g, 256>>>(a); }); report(NAME, ms, g); }
#define RUN34P(NAME, ...) { auto k = k_mc34p<__VA_ARGS__>; int g = occ_grid(k, sms); \
        float ms = time_ms([&] { cudaMemsetAsync(a.counter, 0, 4); k<<<g, 256>>>(a); }); report(NAME, ms, g); }
    RUN34P("4 samples / 3 calls, packed pairs LPB1", 10, 1, 1)
    RUN34P("4 samples / 3 calls, packed pairs LPB1 minb4", 10, 1, 4)
    RUN34P("4 samples / 3 calls, packed pairs LPB1 minb5", 10, 1, 5)
    RUN34P("4 samples / 3 calls, packed pairs philox7", 7, 1, 1)
    RUN34("4 samples / 3 calls LPB1", 10, 1, 1)
    RUN34("4 samples / 3 calls LPB1 minb4", 10, 1, 4)
    RUN34("4 samples / 3 calls LPB1 minb5", 10, 1, 5)
    RUN34("4 samples / 3 calls LPB2", 10, 2, 1)
    RUN34("4 samples / 3 calls LPB1 philox7", 7, 1, 1)
    RUN("packed 1 pair lop LPB1 philox7",   7, 1, 1, 1, 1, 1)
    RUN("packed 2 pairs lop LPB1 philox7",  7, 1, 1, 2, 1, 1)
    {
        int iters = 4096; int g = sms * 8;
        double fl = double(g) * 256 * iters * 8 * 2;
        float ms = time_ms([&] { k_ffma<<<g, 256>>>(a.out, iters, 1.0001f); });
        printf("FFMA imm chain:   %.3f ms  %.1f TFLOP/s\n", ms, fl / ms * 1e-9);
        ms = time_ms([&] { k_ffma2<<<g, 256>>>(a.out, iters, 1.0001f); });
        printf("FFMA2 imm chain:  %.3f ms  %.1f TFLOP/s\n", ms, 2 * fl / ms * 1e-9);
#define MIX(NW, NF, NF2) { float t = time_ms([&] { k_mix<NW, NF, NF2><<<g, 256>>>(a.out, iters, 1.0001f); }); \
        const double cyc = double(t) * 1e-3 * 1.965e9 / (double(iters) * (g * 8 / sms / 4)); \
        printf("mix IMAD.WIDE x%d + FFMA x%d + FFMA2 x%d : %.3f ms  -> %.2f SMSP-cycles per warp-iteration (at 1965 MHz)\n", NW, NF, NF2, t, cyc); }
        MIX(8, 0, 0) MIX(0, 8, 0) MIX(0, 0, 8) MIX(4, 8, 0) MIX(4, 16, 0) MIX(4, 0, 4) MIX(4, 0, 8) MIX(2, 0, 8) MIX(4, 0, 12) MIX(8, 0, 8)
    }
    return 0;
}
