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

namespace viltrum { namespace b200 { namespace device {

constexpr int SCATTER_SMEM_BINS = 8192;     // 32 KiB of privatised bins per CTA

template<class F, int DIM, int DIMBINS, bool EXACT>
__global__ void __launch_bounds__(256)
mc_scatter_kernel(const F f, const vb200_scatter_launch a) {
    __shared__ float s_hist[SCATTER_SMEM_BINS];
    const bool priv = a.nbins_total <= uint64_t(SCATTER_SMEM_BINS);
    if (priv) { for (uint32_t i = threadIdx.x; i < a.nbins_total; i += blockDim.x) s_hist[i] = 0.0f; __syncthreads(); }
    const float factor = float(a.factor);
    float ext[DIM];
#pragma unroll
    for (int i = 0; i < DIM; ++i) ext[i] = a.domain.rmax[i] - a.domain.rmin[i];
    for (uint64_t s = a.sample_begin + uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; s < a.sample_end;
         s += uint64_t(gridDim.x) * blockDim.x) {
        std::array<float, DIM> x;
#pragma unroll
        for (int blk = 0; blk < (DIM + 3) / 4; ++blk) {
            const u32x4 r = philox4x32<10>(u32x4{uint32_t(s), uint32_t(s >> 32), 0xffffffffu, uint32_t(blk)}, a.key0, a.key1);
#pragma unroll
            for (int j = 0; j < 4; ++j) { const int i = blk * 4 + j; if (i < DIM) x[i] = fmaf(pick(r, j), ext[i], a.domain.rmin[i]); }
        }
        uint64_t lin = 0, prod = 1;
#pragma unroll
        for (int i = 0; i < DIMBINS; ++i) {
            const float t = float(a.domain.res[i]) * (x[i] - a.domain.rmin[i]) / ext[i];      // monte-carlo.h:57
            uint64_t p = uint64_t(t);
            if (p >= a.domain.res[i]) p = a.domain.res[i] - 1;      // u01 < 1, so only rounding can get here
            lin += p * prod; prod *= a.domain.res[i];
        }
        const float v = f(x) * factor;
        if (priv) atomicAdd(&s_hist[lin], v); else atomicAdd(&a.out[lin], v);
    }
    if (priv) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < a.nbins_total; i += blockDim.x) { const float v = s_hist[i]; if (v != 0.0f) atomicAdd(&a.out[i], v); }
    }
}

}}} // namespace viltrum::b200::device
