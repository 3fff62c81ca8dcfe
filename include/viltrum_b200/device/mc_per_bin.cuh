// K1 — per-bin / per-stratum Monte Carlo sample kernel, replacing the reference's
//   MonteCarloPerBinParallel::integrate(Range)        src/monte-carlo/monte-carlo-per-bin-parallel.h:41-71
//   IntegratorPerBinParallel + MonteCarlo::integrate  src/integrator-per-bin-parallel.h:16-35, src/monte-carlo/monte-carlo.h:39-63
//   for_each(parallel, multidimensional_range(res))   src/foreach.h:43-78
// One fused pass: Philox4x32-10 draws -> affine map into the bin box -> integrand -> in-bin reduction with warp
// shuffles -> scaled, coalesced store into the tensor bin layout (dim 0 fastest).  Nothing but the final bin
// values touches HBM (4 B per bin per `spp` evaluations), so the kernel is bound by FP32 issue, not bandwidth.
//
// Work decomposition: LPB (a power of two, chosen by the driver from spp) lanes of one warp share a bin and stride
// over its samples; a warp therefore owns 32/LPB consecutive bins per step and writes them as one contiguous run.
// The grid is sized to the resident CTA capacity of the chip (148 SMs x occupancy) and warps pull tiles from a
// global counter until the shard is done.
#pragma once
#include <array>
#include <cuda_runtime.h>
#include "../../viltrum_b200.h"
#include "philox.cuh"
#include "f32x2.cuh"
#include <type_traits>
#include <utility>

namespace viltrum { namespace b200 { namespace device {

constexpr int MC_THREADS = 256;

// bin position (tensor order, dim 0 fastest) -> per-dimension index; reference src/tensor.h:17-23
template<int DIMBINS>
__device__ __forceinline__ void unflatten_bin(uint64_t bin, const vb200_domain& dom, uint32_t (&pos)[VB200_MAX_DIMBINS]) {
    if ((bin >> 32) == 0) {       // 32-bit division: the 64-bit one is a ~100-instruction subroutine
        uint32_t b = uint32_t(bin);
#pragma unroll
        for (int i = 0; i < DIMBINS; ++i) { const uint32_t r = uint32_t(dom.res[i]); pos[i] = b % r; b /= r; }
    } else {
#pragma unroll
        for (int i = 0; i < DIMBINS; ++i) { const uint64_t r = dom.res[i]; pos[i] = uint32_t(bin % r); bin /= r; }
    }
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
            const float drange = dom.drange[i];      // (max-min)/Float(res), computed once on the host (same IEEE division)
            a = __fadd_rn(dom.rmin[i], __fmul_rn(float(pos[i]), drange));
            b = __fadd_rn(dom.rmin[i], __fmul_rn(float(pos[i] + 1u), drange));
        }
        lo[i] = a; ext[i] = b - a;
        volume = __fmul_rn(volume, ext[i]);          // Range::_volume, float product in dimension order (range.h:21-25)
    }
}

// End of a warp's tile on the end-to-end path (vb200_chunk_signal): the lanes' bin stores went to host-mapped memory; order them
// before the chunk counter at system scope, and let the warp that completes a chunk raise the chunk's host flag.
__device__ __forceinline__ void signal_tile_done(const vb200_chunk_signal& c, uint64_t tile, uint64_t ntiles, uint32_t lane) {
    if (!c.enabled) return;
    __syncwarp();                                   // the storing lanes' writes happen-before lane 0's fence
    if (lane == 0) {
        __threadfence_system();
        const uint64_t chunk = tile >> c.chunk_shift;
        const uint64_t first = chunk << c.chunk_shift, per = 1ull << c.chunk_shift;
        const uint32_t count = uint32_t(ntiles - first < per ? ntiles - first : per);
        if (atomicAdd(&c.done[chunk], 1u) + 1u == count) {
            __threadfence_system();
            *reinterpret_cast<volatile uint32_t*>(&c.flag[chunk]) = c.epoch;
        }
    }
}

// EXACT only tags the instantiation (the same template is compiled twice into the library, once in a TU built
// with --fmad=false); it keeps the two sets of kernel symbols apart.
//
// Design notes (measured on B200, profiles/exp/mc_variants.cu, profiles/exp/mc_packed.cu, profiles/exp/pipes.cu):
//   * warp-autonomous tiles: a warp owns G = 32/LPB consecutive bins per step and never meets a CTA barrier; tiles are handed out
//     through one global atomic per warp and tile (a.tile_counter), so the tail is one tile long instead of one static share;
//   * the 32x32->64 multiplies of Philox are the expensive instructions of this kernel — IMAD.WIDE has a reciprocal throughput of
//     5.1 cycles per warp and SMSP and does not overlap with FFMA (profiles/pipes_r1.txt) — so no generated bit is thrown away:
//     samples are drawn in GROUPS OF FOUR from THREE Philox calls per block of four dimensions.  Sample j < 3 of a group takes the
//     top 24 bits of every word of call j, sample 3 is assembled from the three low bytes (two PRMT per coordinate):
//     24 bits x 4 coordinates x 4 samples = 384 bits = 3 x 128.  Counter = (bin lo, bin hi, group, call + 4*block), key = seed;
//   * four samples are in flight per lane (independent Philox + integrand chains); functors that are generic over their scalar
//     type are evaluated as two packed pairs (f32x2.cuh, FFMA2): half the issue slots for the FP32 work;
//   * the u32 -> [0,1) scaling (2^-24) is folded into the bin extent, so a coordinate costs SHF/PRMT + I2FP + FFMA.
template<class F, int DIM, class = void> struct has_pair_eval : std::false_type {};
template<class F, int DIM>
struct has_pair_eval<F, DIM, std::void_t<decltype(std::declval<const F&>()(std::declval<const std::array<f32x2, DIM>&>()))>>
    : std::is_same<decltype(std::declval<const F&>()(std::declval<const std::array<f32x2, DIM>&>())), f32x2> {};

constexpr int MC_GROUP = 4;                 // samples per draw group
template<int DIM> struct GroupDraws {
    static constexpr int NB = (DIM + 3) / 4;      // blocks of four dimensions
    u32x4 r[3][NB];
    __device__ __forceinline__ void draw(uint32_t b0, uint32_t b1, uint32_t group, uint32_t k0, uint32_t k1, int calls) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int blk = 0; blk < NB; ++blk)
                if (j < calls) r[j][blk] = philox4x32<10>(u32x4{b0, b1, group, uint32_t(j + 4 * blk)}, k0, k1);
    }
    __device__ __forceinline__ static uint32_t word(const u32x4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
    // 24-bit integer of coordinate i of sample j (as float: exact)
    __device__ __forceinline__ float coord(int j, int i) const {
        const int blk = i >> 2, w = i & 3;
        if (j < 3) return float(word(r[j][blk], w) >> 8);
        const uint32_t p = word(r[0][blk], w), q = word(r[1][blk], w), t = word(r[2][blk], w);
        return float(__byte_perm(__byte_perm(t, q, 0x7740), p, 0x7410) & 0x00ffffffu);       // (p.b0 << 16) | (q.b0 << 8) | t.b0
    }
};

template<class F, int DIM>
__device__ __forceinline__ float mc_eval_one(const F& f, const GroupDraws<DIM>& d, int j, const float (&lo)[DIM], const float (&ext24)[DIM]) {
    std::array<float, DIM> x;
#pragma unroll
    for (int i = 0; i < DIM; ++i) x[i] = fmaf(d.coord(j, i), ext24[i], lo[i]);     // u*(b-a)+a as std::uniform_real_distribution, u = n*2^-24 in [0,1)
    return f(x);
}
template<class F, int DIM>
__device__ __forceinline__ f32x2 mc_eval_pair(const F& f, const GroupDraws<DIM>& d, int j0, const float (&lo)[DIM], const float (&ext24)[DIM]) {
    std::array<f32x2, DIM> x;
#pragma unroll
    for (int i = 0; i < DIM; ++i) x[i] = mad(f32x2::pack(d.coord(j0, i), d.coord(j0 + 1, i)), f32x2(ext24[i]), f32x2(lo[i]));
    return f(x);
}

template<class F, int DIM, int DIMBINS, bool MOMENTS, bool EXACT>
__global__ void __launch_bounds__(MC_THREADS)
mc_per_bin_kernel(const F f, const vb200_mc_launch a) {
    constexpr bool PAIRS = !EXACT && has_pair_eval<F, DIM>::value;
    const uint32_t LPB = a.lanes_per_bin;     // power of two <= 32: the lanes of a bin sit in one warp
    const uint32_t G = 32u / LPB;             // bins per warp step
    const uint32_t lane = threadIdx.x & 31u, sub = lane % LPB, grp = lane / LPB;
    const uint64_t nshard = a.bin_end - a.bin_begin;
    const uint64_t ntiles = (nshard + G - 1) / G;
    const uint32_t full_groups = a.spp / MC_GROUP, rest = a.spp % MC_GROUP;      // the lanes of a bin stride over its sample groups

    uint64_t tile = 0;
    if (lane == 0) tile = atomicAdd(a.tile_counter, 1ull);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    while (tile < ntiles) {
        const uint64_t bin = a.bin_begin + tile * G + grp;
        const bool live = bin < a.bin_end;
        float sum = 0.0f, sum2 = 0.0f, volume = 1.0f;
        if (live) {
            float lo[DIM], ext[DIM];
            bin_box<DIM, DIMBINS>(a.domain, bin, lo, ext, volume);
#pragma unroll
            for (int i = 0; i < DIM; ++i) ext[i] *= 5.9604644775390625e-08f;
            const uint32_t b0 = uint32_t(bin), b1 = uint32_t(bin >> 32);
            GroupDraws<DIM> d;
            if constexpr (PAIRS) {
                f32x2 acc0(0.0f), acc1(0.0f), sq0(0.0f), sq1(0.0f);
                for (uint32_t g = sub; g < full_groups; g += LPB) {
                    d.draw(b0, b1, g, a.key0, a.key1, 3);
                    const f32x2 v0 = mc_eval_pair<F, DIM>(f, d, 0, lo, ext), v1 = mc_eval_pair<F, DIM>(f, d, 2, lo, ext);
                    acc0 += v0; acc1 += v1;
                    if (MOMENTS) { sq0 = mad(v0, v0, sq0); sq1 = mad(v1, v1, sq1); }
                }
                sum = (acc0.lo() + acc0.hi()) + (acc1.lo() + acc1.hi());
                if (MOMENTS) sum2 = (sq0.lo() + sq0.hi()) + (sq1.lo() + sq1.hi());
            } else {
                float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f, q0 = 0.0f, q1 = 0.0f, q2 = 0.0f, q3 = 0.0f;
                for (uint32_t g = sub; g < full_groups; g += LPB) {
                    d.draw(b0, b1, g, a.key0, a.key1, 3);
                    const float v0 = mc_eval_one<F, DIM>(f, d, 0, lo, ext), v1 = mc_eval_one<F, DIM>(f, d, 1, lo, ext);
                    const float v2 = mc_eval_one<F, DIM>(f, d, 2, lo, ext), v3 = mc_eval_one<F, DIM>(f, d, 3, lo, ext);
                    s0 += v0; s1 += v1; s2 += v2; s3 += v3;
                    if (MOMENTS) { q0 = fmaf(v0, v0, q0); q1 = fmaf(v1, v1, q1); q2 = fmaf(v2, v2, q2); q3 = fmaf(v3, v3, q3); }
                }
                sum = (s0 + s1) + (s2 + s3);
                if (MOMENTS) sum2 = (q0 + q1) + (q2 + q3);
            }
            if (rest != 0u && sub == full_groups % LPB) {      // the last, partial group: one call per sample
                d.draw(b0, b1, full_groups, a.key0, a.key1, int(rest));
                for (uint32_t j = 0; j < rest; ++j) {
                    const float v = j == 0 ? mc_eval_one<F, DIM>(f, d, 0, lo, ext) : j == 1 ? mc_eval_one<F, DIM>(f, d, 1, lo, ext) : mc_eval_one<F, DIM>(f, d, 2, lo, ext);
                    sum += v;
                    if (MOMENTS) sum2 = fmaf(v, v, sum2);
                }
            }
        }
        // in-bin reduction across the LPB lanes that share the bin (all inside one warp)
        for (uint32_t off = LPB >> 1; off > 0; off >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, off);
            if (MOMENTS) sum2 += __shfl_xor_sync(0xffffffffu, sum2, off);
        }
        if (live && sub == 0) {       // G consecutive bins per warp: one contiguous 4*G-byte store
            float v;
            if (a.flavor == VB200_PER_BIN_MC) {
                // sol = sum f * (vol(bin box)/spp) ; bins = double(nbins)*sol   (monte-carlo.h:43-45,59; integrator-per-bin-parallel.h:33)
                const float sol = float(double(sum) * (double(volume) / double(a.spp)));
                v = float(double(a.nbins_total) * double(sol));
            } else {
                v = float(double(sum) * a.factor);                              // monte-carlo-per-bin-parallel.h:45,68
            }
            a.out[bin] = a.accumulate ? float(double(a.out[bin]) + double(v)) : v;   // '+=' vs '=' 
            if (MOMENTS) {
                if (a.sum_f)  a.sum_f[bin - a.bin_begin]  = sum;
                if (a.sum_f2) a.sum_f2[bin - a.bin_begin] = sum2;
            }
        }
        signal_tile_done(a.signal, tile, ntiles, lane);
        if (lane == 0) tile = atomicAdd(a.tile_counter, 1ull);
        tile = __shfl_sync(0xffffffffu, tile, 0);
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
// T = float, or double for VB200_INTEGRAND_F64 integrands.
template<class F, int DIM, class T, bool EXACT>
__global__ void __launch_bounds__(256)
eval_points_kernel(const F f, const vb200_eval_launch a) {
    const T* __restrict__ points = static_cast<const T*>(a.points);
    T* __restrict__ values = static_cast<T*>(a.values);
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n; i += uint64_t(gridDim.x) * blockDim.x) {
        std::array<T, DIM> x;
#pragma unroll
        for (int d = 0; d < DIM; ++d) x[d] = points[uint64_t(d) * a.n + i];
        values[i] = f(x);
    }
}

}}} // namespace viltrum::b200::device
