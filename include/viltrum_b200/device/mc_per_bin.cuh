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
#include "threefry.cuh"
#include "xoshiro.cuh"
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
            c.done[chunk] = 0u;                     // self-resetting: nobody else touches this chunk's counter in this launch
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
//     samples are drawn in GROUPS OF EIGHT whose Philox words are cut into 24-bit and (binned dimensions of fine grids) 16-bit
//     coordinate fields, see GroupDraws.  Counter = (bin lo, bin hi, group, call), key = seed;
//   * functors that are generic over their scalar type are evaluated as packed pairs (f32x2.cuh, FFMA2): half the issue slots
//     for the FP32 work; two pairs (or four scalar samples) are in flight per lane;
//   * the integer -> [0,1) scaling (2^-24 / 2^-16) is folded into the bin extent, so a coordinate costs SHF/LOP/PRMT + I2FP + FFMA.
template<class F, int DIM, class = void> struct has_pair_eval : std::false_type {};
template<class F, int DIM>
struct has_pair_eval<F, DIM, std::void_t<decltype(std::declval<const F&>()(std::declval<const std::array<f32x2, DIM>&>()))>>
    : std::is_same<decltype(std::declval<const F&>()(std::declval<const std::array<f32x2, DIM>&>())), f32x2> {};

constexpr int MC_GROUP = 8;                 // samples per draw group
#ifndef VB200_MC_ROUNDS
#define VB200_MC_ROUNDS 10                  // Philox4x32-10; experiment builds (profiles/exp) measure 7, the product ships 10
#endif
#ifndef VB200_MC_TF_NUM
#define VB200_MC_TF_NUM 0                   // share of a group's generator calls drawn from Threefry4x32: NUM/DEN, rounded
#endif
#ifndef VB200_MC_TF_DEN
#define VB200_MC_TF_DEN 5
#endif
#ifndef VB200_MC_TF_ROUNDS
#define VB200_MC_TF_ROUNDS 12
#endif
// One draw group = 8 samples of one bin.  The group's random words w[0..4*CALLS) are the outputs of CALLS Philox calls with the
// counters (bin lo, bin hi, group, call), and every word is cut into coordinate fields so that no generated bit is thrown away:
//   * 24-bit fields (the reference's generate_canonical<float,24> lattice): three words give four fields — the top 24 bits of
//     each word and a fourth assembled from the three low bytes (two PRMT);
//   * 16-bit fields for the first NARROW dimensions: two per word.  The driver picks NARROW = dimbins only when every binned
//     dimension has >= 256 bins, so that the sample lattice along such a dimension (resolution x 2^16 points over the range)
//     is at least as fine as the reference's (2^24 points over the bin, which float rounding of lo + u*(hi-lo) coarsens to
//     ~2^24 over the range anyway); NARROW = 0 otherwise.
// Words per group: 4*NARROW + 6*(DIM-NARROW).  4-D integrand over a 2-D bin grid: 20 words = 5 calls per 8 samples
// (all-24-bit: 24 words = 6 calls, the round-1b "4 samples per 3 calls").
enum { MC_RNG_XOSHIRO = VB200_RNG_XOSHIRO, MC_RNG_PHILOX = VB200_RNG_PHILOX };     // template argument = vb200_mc_launch.rng
template<int DIM, int NARROW, int RNG = MC_RNG_PHILOX> struct GroupDraws {      // (the scatter kernel keeps the stateless Philox default)
    static constexpr int WIDE = DIM - NARROW;
    static constexpr int W16 = 4 * NARROW;          // first W16 words: 16-bit fields
    static constexpr int W24 = 6 * WIDE;            // then W24 words: 24-bit fields, in triples
    static constexpr int CALLS = (W16 + W24 + 3) / 4;
    uint32_t w[4 * CALLS];
    // RNG = MC_RNG_PHILOX: stateless, the words of group g are Philox4x32-10(key = seed, counter = (bin lo, bin hi, g, call)).
    //   Experiment knob (profiles/exp/k1_mix.cu, profiles/k1_rng_r2.txt): the LAST TF of the CALLS calls can come from Threefry4x32-12
    //   (ALU pipe) instead; measured no faster than pure Philox, so the product builds with TF = 0.
    // RNG = MC_RNG_XOSHIRO: one xoshiro128++ stream per (bin, lane sub-stream), seeded by begin() with
    //   Philox4x32-10(key = seed, counter = (bin lo, bin hi, sub, 'strm')); draw() takes the next W16 + W24 words of the stream.
    static constexpr int TF = (CALLS * VB200_MC_TF_NUM + VB200_MC_TF_DEN / 2) / VB200_MC_TF_DEN;
    uint32_t b0, b1, k0, k1;
    Xoshiro128pp x;
    __device__ __forceinline__ void begin(uint32_t bin_lo, uint32_t bin_hi, uint32_t sub, uint32_t key0, uint32_t key1) {
        b0 = bin_lo; b1 = bin_hi; k0 = key0; k1 = key1;
        if constexpr (RNG == MC_RNG_XOSHIRO) x.seed(philox4x32<10>(u32x4{bin_lo, bin_hi, sub, 0x7374726du}, key0, key1));
    }
    __device__ __forceinline__ void draw(uint32_t group) {
        if constexpr (RNG == MC_RNG_XOSHIRO) {
#pragma unroll
            for (int i = 0; i < W16 + W24; ++i) w[i] = x.next();
        } else {
            const ThreefryKeys tk = threefry_key_schedule(k0, k1, 0x76696c74u, 0x72756d21u);      // key words 2,3: "vilt" "rum!"
#pragma unroll
            for (int c = 0; c < CALLS; ++c) {
                u32x4 r;
                if (c >= CALLS - TF) r = threefry4x32<VB200_MC_TF_ROUNDS>(u32x4{b0, b1, group, uint32_t(c)}, tk);
                else r = philox4x32<VB200_MC_ROUNDS>(u32x4{b0, b1, group, uint32_t(c)}, k0, k1);
                w[4 * c] = r.x; w[4 * c + 1] = r.y; w[4 * c + 2] = r.z; w[4 * c + 3] = r.w;
            }
        }
    }
    // integer n of coordinate i of sample j as a float (exact): 24-bit fields give n, 16-bit fields (i < NARROW) give 2^23 + n,
    // built straight from the bits with one LOP3/PRMT (cvt.f32.u16 becomes I2F.U16 on the quarter-rate XU pipe); the kernel folds
    // the 2^23 into the bin's lower corner (bias()).  Samples 2p and 2p+1 share the word of a 16-bit field.
    __device__ __forceinline__ static constexpr float bias(int i) { return i < NARROW ? 8388608.0f : 0.0f; }
    __device__ __forceinline__ float coord(int j, int i) const {
        if (i < NARROW) {
            const uint32_t v = w[(j >> 1) * NARROW + i];
            return __uint_as_float((j & 1) ? __byte_perm(v, 0x4b000000u, 0x7632) : ((v & 0xffffu) | 0x4b000000u));
        }
        const int m = j * WIDE + (i - NARROW), base = W16 + 3 * (m >> 2), q = m & 3;
        if (q < 3) return float(w[base + q] >> 8);
        return float(__byte_perm(__byte_perm(w[base + 2], w[base + 1], 0x7740), w[base], 0x7410) & 0x00ffffffu);   // (w0.b0 << 16) | (w1.b0 << 8) | w2.b0
    }
};

template<class F, int DIM, class Draws>
__device__ __forceinline__ float mc_eval_one(const F& f, const Draws& d, int j, const float (&lo)[DIM], const float (&exts)[DIM]) {
    std::array<float, DIM> x;
#pragma unroll
    for (int i = 0; i < DIM; ++i) x[i] = fmaf(d.coord(j, i), exts[i], lo[i]);     // u*(b-a)+a as std::uniform_real_distribution, u = n*2^-bits in [0,1)
    return f(x);
}
template<class F, int DIM, class Draws>
__device__ __forceinline__ f32x2 mc_eval_pair(const F& f, const Draws& d, int j0, const float (&lo)[DIM], const float (&exts)[DIM]) {
    std::array<f32x2, DIM> x;
#pragma unroll
    for (int i = 0; i < DIM; ++i) x[i] = mad(f32x2::pack(d.coord(j0, i), d.coord(j0 + 1, i)), f32x2(exts[i]), f32x2(lo[i]));
    return f(x);
}

#ifndef VB200_MC_PIPELINE
#define VB200_MC_PIPELINE 0
#endif
#ifndef VB200_MC_CHAINS
#define VB200_MC_CHAINS 2
#endif
#ifndef VB200_MC_MINB
#define VB200_MC_MINB 3                     // 3 CTAs/SM (<= 85 registers: room for heavier user functors and the Fubini adapters); 4 CTAs/SM at 64 registers is within 1 % (profiles/k1_rng_r2.txt)
#endif
template<class F, int DIM, int DIMBINS, bool MOMENTS, bool EXACT, bool NARROW, int RNG = MC_RNG_PHILOX>
__global__ void __launch_bounds__(MC_THREADS, VB200_MC_MINB)
mc_per_bin_kernel(const F f, const vb200_mc_launch a) {
    constexpr bool PAIRS = !EXACT && has_pair_eval<F, DIM>::value;
    constexpr int NB = NARROW ? DIMBINS : 0;  // dimensions drawn as 16-bit fields
    const uint32_t LPB = a.lanes_per_bin;     // power of two <= 32: the lanes of a bin sit in one warp
    const uint32_t G = 32u / LPB;             // bins per warp step
    const uint32_t lane = threadIdx.x & 31u, sub = lane % LPB, grp = lane / LPB;
    const uint64_t nshard = a.bin_end - a.bin_begin;
    const uint64_t ntiles = (nshard + G - 1) / G;
    const uint32_t full_groups = a.spp / MC_GROUP, rest = a.spp % MC_GROUP;      // the lanes of a bin stride over its sample groups

    // dynamic tile scheduler: tickets from a.tile_counter[0].  The counters reset themselves: a.tile_counter[1] counts the warps that have
    // drawn their last ticket, and the last of them zeroes both words — the driver never has to clear them between launches.  (Drawing the
    // next ticket ahead of the current tile, to hide the atomic's round trip, was measured and dropped: two more live registers cost the
    // loop more than the hidden latency gave back, profiles/k1_rng_r2.txt.)
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
            for (int i = 0; i < DIM; ++i) {       // 2^-16 / 2^-24 folded into the extent, the 16-bit fields' 2^23 into the lower corner
                ext[i] *= (i < NB ? 1.52587890625e-05f : 5.9604644775390625e-08f);
                lo[i] = fmaf(-GroupDraws<DIM, NB>::bias(i), ext[i], lo[i]);
            }
            const uint32_t b0 = uint32_t(bin), b1 = uint32_t(bin >> 32);
            GroupDraws<DIM, NB, RNG> d;
            d.begin(b0, b1, sub, a.key0, a.key1);
            if constexpr (PAIRS) {
                f32x2 acc0(0.0f), acc1(0.0f), sq0(0.0f), sq1(0.0f);
#if VB200_MC_PIPELINE
                // software pipeline: the words of the NEXT group are drawn while the current group is evaluated — two independent
                // instruction streams in one basic block (generator: ALU pipe, integrand: FMA pipe), which ptxas interleaves
                if (sub < full_groups) d.draw(sub);
                for (uint32_t g = sub; g < full_groups; g += LPB) {
                    const GroupDraws<DIM, NB, RNG> cur = d;
                    if (g + LPB < full_groups) d.draw(g + LPB);
#else
                for (uint32_t g = sub; g < full_groups; g += LPB) {
                    d.draw(g);
                    const GroupDraws<DIM, NB, RNG>& cur = d;
#endif
#if VB200_MC_CHAINS == 4
                    const f32x2 v0 = mc_eval_pair<F, DIM>(f, cur, 0, lo, ext), v1 = mc_eval_pair<F, DIM>(f, cur, 2, lo, ext),
                                v2 = mc_eval_pair<F, DIM>(f, cur, 4, lo, ext), v3 = mc_eval_pair<F, DIM>(f, cur, 6, lo, ext);
                    acc0 += v0; acc1 += v1; acc0 += v2; acc1 += v3;
                    if (MOMENTS) { sq0 = mad(v0, v0, sq0); sq1 = mad(v1, v1, sq1); sq0 = mad(v2, v2, sq0); sq1 = mad(v3, v3, sq1); }
#else
                    const f32x2 v0 = mc_eval_pair<F, DIM>(f, cur, 0, lo, ext), v1 = mc_eval_pair<F, DIM>(f, cur, 2, lo, ext);
                    acc0 += v0; acc1 += v1;
                    if (MOMENTS) { sq0 = mad(v0, v0, sq0); sq1 = mad(v1, v1, sq1); }
                    const f32x2 v2 = mc_eval_pair<F, DIM>(f, cur, 4, lo, ext), v3 = mc_eval_pair<F, DIM>(f, cur, 6, lo, ext);
                    acc0 += v2; acc1 += v3;
                    if (MOMENTS) { sq0 = mad(v2, v2, sq0); sq1 = mad(v3, v3, sq1); }
#endif
                }
                sum = (acc0.lo() + acc0.hi()) + (acc1.lo() + acc1.hi());
                if (MOMENTS) sum2 = (sq0.lo() + sq0.hi()) + (sq1.lo() + sq1.hi());
            } else {
                float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f, q0 = 0.0f, q1 = 0.0f, q2 = 0.0f, q3 = 0.0f;
                for (uint32_t g = sub; g < full_groups; g += LPB) {
                    d.draw(g);
#pragma unroll
                    for (int h = 0; h < MC_GROUP; h += 4) {
                        const float v0 = mc_eval_one<F, DIM>(f, d, h, lo, ext), v1 = mc_eval_one<F, DIM>(f, d, h + 1, lo, ext);
                        const float v2 = mc_eval_one<F, DIM>(f, d, h + 2, lo, ext), v3 = mc_eval_one<F, DIM>(f, d, h + 3, lo, ext);
                        s0 += v0; s1 += v1; s2 += v2; s3 += v3;
                        if (MOMENTS) { q0 = fmaf(v0, v0, q0); q1 = fmaf(v1, v1, q1); q2 = fmaf(v2, v2, q2); q3 = fmaf(v3, v3, q3); }
                    }
                }
                sum = (s0 + s1) + (s2 + s3);
                if (MOMENTS) sum2 = (q0 + q1) + (q2 + q3);
            }
            if (rest != 0u && sub == full_groups % LPB) {      // the last, partial group: its first `rest` samples
                d.draw(full_groups);
#pragma unroll
                for (int j = 0; j < MC_GROUP - 1; ++j) {
                    if (uint32_t(j) < rest) {
                        const float v = mc_eval_one<F, DIM>(f, d, j, lo, ext);
                        sum += v;
                        if (MOMENTS) sum2 = fmaf(v, v, sum2);
                    }
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
    if (lane == 0) {
        const unsigned long long finished = atomicAdd(a.tile_counter + 1, 1ull) + 1ull;
        if (finished == uint64_t(gridDim.x) * (MC_THREADS / 32)) { a.tile_counter[0] = 0ull; a.tile_counter[1] = 0ull; }
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
