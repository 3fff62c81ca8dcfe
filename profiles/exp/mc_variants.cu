// Experiment harness (not product): times variants of the per-bin MC kernel on the C2 workload so that kernel-design
// choices are measured, not guessed.  nvcc -O3 -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <array>
#include <cuda_runtime.h>
#include <viltrum_b200/device/philox.cuh>
#include <viltrum_b200/device/mc_per_bin.cuh>
#include "../../viltrum_b200/csrc/builtin_integrands.cuh"

using namespace viltrum::b200;
using F = builtin::Shade4<64>;

struct Args { uint32_t res0, res1; uint32_t spp; uint32_t k0, k1; float* out; uint64_t nbins; unsigned* counter; };

template<int ROUNDS, bool FUSED>
__device__ __forceinline__ void draw4(uint32_t b0, uint32_t s, uint32_t k0, uint32_t k1, const float (&lo)[4], const float (&ext)[4], std::array<float,4>& x) {
    const u32x4 r = philox4x32<ROUNDS>(u32x4{b0, 0u, s, 0u}, k0, k1);
    if (FUSED) {   // ext pre-scaled by 2^-24: x = float(u>>8)*ext24 + lo  (identical bits: power-of-two scaling is exact)
        x[0] = fmaf(float(r.x >> 8), ext[0], lo[0]); x[1] = fmaf(float(r.y >> 8), ext[1], lo[1]);
        x[2] = fmaf(float(r.z >> 8), ext[2], lo[2]); x[3] = fmaf(float(r.w >> 8), ext[3], lo[3]);
    } else {
        x[0] = fmaf(u01(r.x), ext[0], lo[0]); x[1] = fmaf(u01(r.y), ext[1], lo[1]);
        x[2] = fmaf(u01(r.z), ext[2], lo[2]); x[3] = fmaf(u01(r.w), ext[3], lo[3]);
    }
}

__device__ __forceinline__ void box(const Args& a, uint32_t bin, bool fused, float (&lo)[4], float (&ext)[4]) {
    const uint32_t p0 = bin % a.res0, p1 = bin / a.res0;
    const float d0 = 1.0f / float(a.res0), d1 = 1.0f / float(a.res1);
    lo[0] = float(p0) * d0; ext[0] = float(p0 + 1) * d0 - lo[0];
    lo[1] = float(p1) * d1; ext[1] = float(p1 + 1) * d1 - lo[1];
    lo[2] = 0; ext[2] = 1; lo[3] = 0; ext[3] = 1;
    if (fused) for (int i = 0; i < 4; ++i) ext[i] *= 5.9604644775390625e-08f;
}

// warp-autonomous: each warp owns 32/LPB bins per step, no CTA barrier; MINB = min blocks/SM for launch bounds
template<int LPB, int ROUNDS, bool FUSED, int ILP, bool DYNAMIC, int MINB>
__global__ void __launch_bounds__(256, MINB) k_warp(const F f, const Args a) {
    constexpr int G = 32 / LPB;
    const uint32_t lane = threadIdx.x & 31, sub = lane % LPB, grp = lane / LPB;
    const uint32_t warp = (blockIdx.x * 256 + threadIdx.x) >> 5, nwarps = (gridDim.x * 256) >> 5;
    const uint32_t ntiles = uint32_t(a.nbins / G);
    uint32_t tile = warp;
    if (DYNAMIC) { if (lane == 0) tile = atomicAdd(a.counter, 1u); tile = __shfl_sync(0xffffffffu, tile, 0); }
    while (tile < ntiles) {
        const uint32_t bin = tile * G + grp;
        float lo[4], ext[4]; box(a, bin, FUSED, lo, ext);
        float sum = 0.f;
        if (ILP == 1) {
            for (uint32_t s = sub; s < a.spp; s += LPB) { std::array<float,4> x; draw4<ROUNDS, FUSED>(bin, s, a.k0, a.k1, lo, ext, x); sum += f(x); }
        } else {
            float sum2 = 0.f;
            for (uint32_t s = sub; s < a.spp; s += 2 * LPB) {
                std::array<float,4> x, y; draw4<ROUNDS, FUSED>(bin, s, a.k0, a.k1, lo, ext, x); draw4<ROUNDS, FUSED>(bin, s + LPB, a.k0, a.k1, lo, ext, y);
                sum += f(x); sum2 += f(y);
            }
            sum += sum2;
        }
#pragma unroll
        for (int off = LPB / 2; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if (sub == 0) a.out[bin] = sum * (1.0f / float(a.spp));
        if (DYNAMIC) { if (lane == 0) tile = atomicAdd(a.counter, 1u); tile = __shfl_sync(0xffffffffu, tile, 0); }
        else tile += nwarps;
    }
}

// CTA tile + barrier (the shipped round-1a design)
template<int LPB, int ROUNDS>
__global__ void __launch_bounds__(256) k_cta(const F f, const Args a) {
    constexpr int BPT = 256 / LPB;
    __shared__ float s_val[2][BPT];
    const uint32_t tid = threadIdx.x, slot = tid / LPB, sub = tid % LPB;
    const uint32_t ntiles = uint32_t(a.nbins / BPT);
    int buf = 0;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
        const uint32_t bin = tile * BPT + slot;
        float lo[4], ext[4]; box(a, bin, false, lo, ext);
        float sum = 0.f;
        for (uint32_t s = sub; s < a.spp; s += LPB) { std::array<float,4> x; draw4<ROUNDS, false>(bin, s, a.k0, a.k1, lo, ext, x); sum += f(x); }
#pragma unroll
        for (int off = LPB / 2; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if (sub == 0) s_val[buf][slot] = sum * (1.0f / float(a.spp));
        __syncthreads();
        if (tid < BPT) a.out[tile * BPT + tid] = s_val[buf][tid];
    }
}

// pure FMA-chain peak: 8 independent chains per thread, immediate-free 3-register FFMA
__global__ void __launch_bounds__(256) k_fma_peak(float* out, int iters, float a, float b) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = float(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], a, b);
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += v[i];
    if (s == 12345.678f) out[0] = s;
}
// same with an immediate addend (the Horner form)
__global__ void __launch_bounds__(256) k_fma_peak_imm(float* out, int iters, float a) {
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

template<class K> int occ_grid(K k, int sms) { int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 256, 0); return occ * sms; }

template<class L> float time_ms(L launch, int reps = 10) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch();
    cudaDeviceSynchronize();
    float best = 1e30f, tot = 0;
    for (int i = 0; i < reps; ++i) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms); tot += ms; }
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
    F f;
#define RUN_WARP(NAME, ...) { auto k = k_warp<__VA_ARGS__>; int g = occ_grid(k, sms); \
        float ms = time_ms([&] { cudaMemsetAsync(a.counter, 0, 4); k<<<g, 256>>>(f, a); }); report(NAME, ms, g); }
    { auto k = k_cta<8, 10>; int g = occ_grid(k, sms); float ms = time_ms([&] { k<<<g, 256>>>(f, a); }); report("cta-tile LPB8 barrier (r1a)", ms, g); }
    { auto k = k_cta<4, 10>; int g = occ_grid(k, sms); float ms = time_ms([&] { k<<<g, 256>>>(f, a); }); report("cta-tile LPB4 barrier (r1a shipped)", ms, g); }
    RUN_WARP("warp LPB4 static", 4, 10, false, 1, false, 1)
    RUN_WARP("warp LPB4 static fused", 4, 10, true, 1, false, 1)
    RUN_WARP("warp LPB8 static fused", 8, 10, true, 1, false, 1)
    RUN_WARP("warp LPB1 static fused (thread per bin)", 1, 10, true, 1, false, 1)
    RUN_WARP("warp LPB4 dynamic fused", 4, 10, true, 1, true, 1)
    RUN_WARP("warp LPB1 dynamic fused", 1, 10, true, 1, true, 1)
    RUN_WARP("warp LPB4 static fused minb6", 4, 10, true, 1, false, 6)
    RUN_WARP("warp LPB4 dynamic fused minb6", 4, 10, true, 1, true, 6)
    RUN_WARP("warp LPB4 static fused ILP2", 4, 10, true, 2, false, 1)
    RUN_WARP("warp LPB4 dynamic fused ILP2", 4, 10, true, 2, true, 1)
    RUN_WARP("warp LPB4 dynamic fused ILP2 minb4", 4, 10, true, 2, true, 4)
    RUN_WARP("warp LPB4 dynamic fused minb8", 4, 10, true, 1, true, 8)
    RUN_WARP("warp LPB4 dynamic fused philox7", 4, 7, true, 1, true, 1)
    {   // the shipped product kernel, launched exactly as the library does
        vb200_mc_launch L; memset(&L, 0, sizeof(L));
        L.domain.dim = 4; L.domain.dimbins = 2; for (int i = 0; i < 4; ++i) { L.domain.rmin[i] = 0; L.domain.rmax[i] = 1; }
        L.domain.res[0] = 1024; L.domain.res[1] = 1024; L.domain.drange[0] = L.domain.drange[1] = 1.0f / 1024.0f; L.bin_begin = 0; L.bin_end = a.nbins; L.nbins_total = a.nbins; L.spp = 64;
        L.key0 = 1; L.key1 = 2; L.flavor = 0; L.accumulate = 0; L.factor = 1.0 / 64; L.out = a.out;
        unsigned long long* ctr; cudaMalloc(&ctr, 8); L.tile_counter = ctr;
        for (uint32_t lpb : {1u, 2u, 4u, 8u}) {
            L.lanes_per_bin = lpb;
            auto k = device::mc_per_bin_kernel<F, 4, 2, false, false>; int g = occ_grid(k, sms);
            float ms = time_ms([&] { cudaMemsetAsync(ctr, 0, 8); k<<<g, 256>>>(f, L); });
            char name[64]; snprintf(name, sizeof(name), "PRODUCT mc_per_bin_kernel LPB%u", lpb); report(name, ms, g);
        }
    }
    {   // FMA peaks
        int iters = 4096; int g = sms * 8;
        float ms = time_ms([&] { k_fma_peak<<<g, 256>>>(a.out, iters, 1.0001f, 0.5f); });
        double fl = double(g) * 256 * iters * 8 * 2;
        printf("FMA chain peak (3-reg FFMA):  %.3f ms  %.1f TFLOP/s\n", ms, fl / ms * 1e-9);
        ms = time_ms([&] { k_fma_peak_imm<<<g, 256>>>(a.out, iters, 1.0001f); });
        printf("FMA chain peak (imm FFMA):    %.3f ms  %.1f TFLOP/s\n", ms, fl / ms * 1e-9);
    }
    return 0;
}
