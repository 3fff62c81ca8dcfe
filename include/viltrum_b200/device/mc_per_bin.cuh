// K1 — per-bin / per-stratum Monte Carlo sample kernel, replacing the reference's
//   MonteCarloPerBinParallel::integrate(Range)        src/monte-carlo/monte-carlo-per-bin-parallel.h:41-71
//   IntegratorPerBinParallel + MonteCarlo::integrate  src/integrator-per-bin-parallel.h:16-35, src/monte-carlo/monte-carlo.h:39-63
//   for_each(parallel, multidimensional_range(res))   src/foreach.h:43-78
// One fused pass: Philox4x32-10 draws -> affine map into the bin box -> integrand -> in-bin reduction with warp
// shuffles -> scaled, coalesced store into the tensor bin layout (dim 0 fastest).  Nothing but the final bin
// values touches HBM (4 B per bin per `spp` evaluations), so the kernel is bound by FP32 issue, not bandwidth.
//
// Work decomposition: a CTA of 256 threads owns one *tile* of 256/LPB consecutive bins; LPB (a power of two,
// chosen by the driver from spp) lanes of one warp share a bin and stride over its samples.  Tiles are handed out grid-stride over a grid sized
// to the resident CTA capacity of the chip (148 SMs x occupancy), so the tail is a fraction of one tile.
// Each tile's bin values are staged in shared memory and written by the first 256/LPB threads as full,
// contiguous lines.
#pragma once
#include <array>
#include <cuda_runtime.h>
#include "../../viltrum_b200.h"
#include "philox.cuh"

namespace viltrum { namespace b200 { namespace device {

constexpr int MC_THREADS = 256;

// bin position (tensor order, dim 0 fastest) -> per-dimension index; reference src/tensor.h:17-23
template<int DIMBINS>
__device__ __forceinline__ void unflatten_bin(uint64_t bin, const vb200_domain& dom, uint32_t (&pos)[VB200_MAX_DIMBINS]) {
#pragma unroll
    for (int i = 0; i < DIMBINS; ++i) { uint64_t r = dom.res[i]; pos[i] = uint32_t(bin % r); bin /= r; }
}

// Bin sub-box exactly as the reference computes it (monte-carlo-per-bin-parallel.h:45-47,59-61):
//   drange = (max-min)/Float(res);  [min + pos*drange, min + (pos+1)*drange]
template<int DIM, int DIMBINS>
__device__ __forceinline__ void bin_box(const vb200_domain& dom, uint64_t bin, float (&lo)[DIM], float (&ext)[DIM], float& volume) {
    uint32_t pos[VB200_MAX_DIMBINS];
    unflatten_bin<DIMBINS>(bin, dom, pos);
    volume = 1.0f;
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
        float a = dom.rmin[i], b = dom.rmax[i];
        if (i < DIMBINS) {
            float drange = __fdiv_rn(dom.rmax[i] - dom.rmin[i], float(dom.res[i]));
            a = __fadd_rn(dom.rmin[i], __fmul_rn(float(pos[i]), drange));
            b = __fadd_rn(dom.rmin[i], __fmul_rn(float(pos[i] + 1u), drange));
        }
        lo[i] = a; ext[i] = b - a;
        volume = __fmul_rn(volume, ext[i]);          // Range::_volume, float product in dimension order (range.h:21-25)
    }
}

// EXACT only tags the instantiation (the same template is compiled twice into the library, once in a TU built
// with --fmad=false); it keeps the two sets of kernel symbols apart.
template<class F, int DIM, int DIMBINS, bool MOMENTS, bool EXACT>
__global__ void __launch_bounds__(MC_THREADS)
mc_per_bin_kernel(const F f, const vb200_mc_launch a) {
    __shared__ float s_val[2][MC_THREADS];
    __shared__ float s_m1[MOMENTS ? MC_THREADS : 1];
    __shared__ float s_m2[MOMENTS ? MC_THREADS : 1];

    const uint32_t LPB = a.lanes_per_bin;     // power of two <= 32: the lanes of a bin sit in one warp
    const uint32_t BINS_PER_TILE = MC_THREADS / LPB;
    const uint32_t tid = threadIdx.x;
    const uint32_t slot = tid / LPB;          // bin within the tile
    const uint32_t sub = tid % LPB;           // lane within the bin
    const uint64_t nshard = a.bin_end - a.bin_begin;
    const uint64_t ntiles = (nshard + BINS_PER_TILE - 1) / BINS_PER_TILE;
    int buf = 0;

    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
        const uint64_t bin = a.bin_begin + tile * BINS_PER_TILE + slot;
        const bool live = bin < a.bin_end;
        float sum = 0.0f, sum2 = 0.0f, volume = 1.0f;
        if (live) {
            float lo[DIM], ext[DIM];
            bin_box<DIM, DIMBINS>(a.domain, bin, lo, ext, volume);
            const uint32_t b0 = uint32_t(bin), b1 = uint32_t(bin >> 32);
            for (uint32_t s = sub; s < a.spp; s += LPB) {
                std::array<float, DIM> x;
#pragma unroll
                for (int blk = 0; blk < (DIM + 3) / 4; ++blk) {
                    const u32x4 r = philox4x32<10>(u32x4{b0, b1, s, uint32_t(blk)}, a.key0, a.key1);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int i = blk * 4 + j;
                        if (i < DIM) x[i] = fmaf(pick(r, j), ext[i], lo[i]);   // u*(b-a)+a, as uniform_real_distribution
                    }
                }
                const float v = f(x);
                sum += v;
                if (MOMENTS) sum2 = fmaf(v, v, sum2);
            }
        }
        // in-bin reduction across the LPB lanes that share the bin (all inside one warp)
        for (uint32_t off = LPB >> 1; off > 0; off >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, off);
            if (MOMENTS) sum2 += __shfl_xor_sync(0xffffffffu, sum2, off);
        }
        if (sub == 0) {
            float v;
            if (a.flavor == VB200_PER_BIN_MC) {
                // sol = sum f * (vol(bin box)/spp) ; bins = double(nbins)*sol   (monte-carlo.h:43-45,59; integrator-per-bin-parallel.h:33)
                const float sol = float(double(sum) * (double(volume) / double(a.spp)));
                v = float(double(a.nbins_total) * double(sol));
            } else {
                v = float(double(sum) * a.factor);                              // monte-carlo-per-bin-parallel.h:45,68
            }
            s_val[buf][slot] = v;
            if (MOMENTS) { s_m1[slot] = sum; s_m2[slot] = sum2; }
        }
        __syncthreads();
        if (tid < BINS_PER_TILE) {
            const uint64_t ob = a.bin_begin + tile * BINS_PER_TILE + tid;
            if (ob < a.bin_end) {
                const float v = s_val[buf][tid];
                a.out[ob] = a.accumulate ? float(double(a.out[ob]) + double(v)) : v;   // '+=' vs '=' 
                if (MOMENTS) {
                    if (a.sum_f)  a.sum_f[ob - a.bin_begin]  = s_m1[tid];
                    if (a.sum_f2) a.sum_f2[ob - a.bin_begin] = s_m2[tid];
                }
            }
        }
        if (MOMENTS) __syncthreads();     // s_m1/s_m2 are single-buffered
    }
}

// K2 — sample replay: recorded sample points, one thread per bin, the reference's sequential arithmetic.
//   flavor 0: bins(p) += f(x)*factor per sample, float(double(acc)+double(f)*factor)      (monte-carlo-per-bin-parallel.h:68)
//   flavor 1: sol += f(x)*factor_bin ; bins(p) = double(nbins)*sol                        (monte-carlo.h:59, integrator-per-bin-parallel.h:33)
template<class F, int DIM, int DIMBINS, bool EXACT>
__global__ void __launch_bounds__(128)
mc_replay_kernel(const F f, const vb200_replay_launch a) {
    const uint64_t k = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t bin = a.bin_begin + k;
    if (bin >= a.bin_end) return;
    float lo[DIM], ext[DIM], volume;
    bin_box<DIM, DIMBINS>(a.domain, bin, lo, ext, volume);
    const double factor = (a.flavor == VB200_PER_BIN_MC) ? __ddiv_rn(double(volume), double(a.spp)) : a.factor;
    float acc = (a.flavor == VB200_PER_BIN_MC) ? 0.0f : a.out[bin];
    const float* s = a.samples + k * uint64_t(a.spp) * DIM;
    for (uint32_t i = 0; i < a.spp; ++i) {
        std::array<float, DIM> x;
#pragma unroll
        for (int d = 0; d < DIM; ++d) x[d] = s[uint64_t(i) * DIM + d];
        const float v = f(x);
        acc = __double2float_rn(__dadd_rn(double(acc), __dmul_rn(double(v), factor)));
    }
    a.out[bin] = (a.flavor == VB200_PER_BIN_MC) ? __double2float_rn(__dmul_rn(double(a.nbins_total), double(acc))) : acc;
}

// K-eval — values[i] = f(points[:, i]); points are SoA (points[d*n+i]) so loads and the store are coalesced.
// Serves the region generators (fill / batched split evaluation) and the control-variate residual pass.
template<class F, int DIM, bool EXACT>
__global__ void __launch_bounds__(256)
eval_points_kernel(const F f, const vb200_eval_launch a) {
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n; i += uint64_t(gridDim.x) * blockDim.x) {
        std::array<float, DIM> x;
#pragma unroll
        for (int d = 0; d < DIM; ++d) x[d] = a.points[uint64_t(d) * a.n + i];
        a.values[i] = f(x);
    }
}

}}} // namespace viltrum::b200::device
