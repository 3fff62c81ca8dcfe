// K4 — global Monte Carlo with scatter binning, replacing reference MonteCarlo::integrate(Range)
// (src/monte-carlo/monte-carlo.h:39-63): samples are drawn over the WHOLE range, the bin is found from the first
// DIMBINS coordinates (pos[i] = size_t(res[i]*(x[i]-min_i)/(max_i-min_i)), :55-58) and bins(pos) += f(x)*factor.
// Few bins + many samples is the one shape where a bin's samples are split across lanes/CTAs/GPUs: per-CTA
// privatised shared-memory histograms (contention stays on chip), one global atomicAdd per bin per CTA, and —
// across GPUs — the caller's allreduce over the partial grids.
#pragma once
#include <array>
#include <cuda_runtime.h>
#include "../../viltrum_b200.h"
#include "philox.cuh"
#include "mc_per_bin.cuh"

namespace viltrum { namespace b200 { namespace device {

constexpr int SCATTER_SMEM_BINS = 8192;     // 32 KiB of privatised bins per CTA

// Samples are drawn like the per-bin sampler's (mc_per_bin.cuh GroupDraws): in groups of eight whose Philox words — counter
// (group lo, group hi, 0xffffffff, call), group = sample / 8 — are cut into 24-bit coordinate fields, 6*DIM words per group (DIM = 2:
// 3 calls per 8 samples instead of 8), and functors that are generic over their scalar type are evaluated as packed pairs.  A sample's
// value depends on its global index only, so sample ranges drawn by different calls / GPUs add up to the single-call estimate.
template<class F, int DIM, int DIMBINS, bool EXACT>
__global__ void __launch_bounds__(256)
mc_scatter_kernel(const F f, const vb200_scatter_launch a) {
    constexpr bool PAIRS = !EXACT && has_pair_eval<F, DIM>::value;
    __shared__ float s_hist[SCATTER_SMEM_BINS];
    const bool priv = a.nbins_total <= uint64_t(SCATTER_SMEM_BINS);
    if (priv) { for (uint32_t i = threadIdx.x; i < a.nbins_total; i += blockDim.x) s_hist[i] = 0.0f; __syncthreads(); }
    const float factor = float(a.factor);
    float ext[DIM], ext24[DIM], lo[DIM];
#pragma unroll
    for (int i = 0; i < DIM; ++i) { lo[i] = a.domain.rmin[i]; ext[i] = a.domain.rmax[i] - a.domain.rmin[i]; ext24[i] = ext[i] * 5.9604644775390625e-08f; }
    auto deposit = [&] (uint64_t s, const float (&xb)[DIMBINS], float fv) {
        if (s < a.sample_begin || s >= a.sample_end) return;
        uint64_t lin = 0, prod = 1;
#pragma unroll
        for (int i = 0; i < DIMBINS; ++i) {
            const float t = float(a.domain.res[i]) * (xb[i] - lo[i]) / ext[i];      // monte-carlo.h:57
            uint64_t p = uint64_t(t);
            if (p >= a.domain.res[i]) p = a.domain.res[i] - 1;      // u01 < 1, so only rounding can get here
            lin += p * prod; prod *= a.domain.res[i];
        }
        const float v = fv * factor;
        if (priv) atomicAdd(&s_hist[lin], v); else atomicAdd(&a.out[lin], v);
    };
    const uint64_t g_begin = a.sample_begin / MC_GROUP, g_end = (a.sample_end + MC_GROUP - 1) / MC_GROUP;
    for (uint64_t g = g_begin + uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; g < g_end; g += uint64_t(gridDim.x) * blockDim.x) {
        GroupDraws<DIM, 0> d;
        d.begin(uint32_t(g), uint32_t(g >> 32), 0u, a.key0, a.key1);
        d.draw(0xffffffffu);
#pragma unroll
        for (int j = 0; j < MC_GROUP; j += 2) {
            float xa[DIMBINS], xb[DIMBINS], va, vb;
            if constexpr (PAIRS) {
                std::array<f32x2, DIM> x;
#pragma unroll
                for (int i = 0; i < DIM; ++i) x[i] = mad(f32x2::pack(d.coord(j, i), d.coord(j + 1, i)), f32x2(ext24[i]), f32x2(lo[i]));
                const f32x2 v = f(x);
                va = v.lo(); vb = v.hi();
#pragma unroll
                for (int i = 0; i < DIMBINS; ++i) { xa[i] = x[i].lo(); xb[i] = x[i].hi(); }
            } else {
                std::array<float, DIM> x0, x1;
#pragma unroll
                for (int i = 0; i < DIM; ++i) { x0[i] = fmaf(d.coord(j, i), ext24[i], lo[i]); x1[i] = fmaf(d.coord(j + 1, i), ext24[i], lo[i]); }
                va = f(x0); vb = f(x1);
#pragma unroll
                for (int i = 0; i < DIMBINS; ++i) { xa[i] = x0[i]; xb[i] = x1[i]; }
            }
            deposit(g * MC_GROUP + uint64_t(j), xa, va);
            deposit(g * MC_GROUP + uint64_t(j) + 1, xb, vb);
        }
    }
    if (priv) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < a.nbins_total; i += blockDim.x) { const float v = s_hist[i]; if (v != 0.0f) atomicAdd(&a.out[i], v); }
    }
}

}}} // namespace viltrum::b200::device
