// Control variates + residual Monte Carlo (kernels K9-K11 of §2.2), replacing the reference's
//   RegionsIntegratorParallelVarianceReduction::integrate_regions   src/control-variates/regions-integrator-parallel-variance-reduction.h:32-109
//   rr_uniform_region / region_sampling_uniform                     src/control-variates/region-russian-roulette.h:9-28, region-sampling.h:9-20
//   Region::approximation_at -> app_at                              src/newton-cotes/region.h:86-112
//   cv_optimize_weight::Accumulator                                 src/control-variates/weight-strategy.h:40-110
// Pipeline per shard of bins (all on the device; the integrand is reached through the eval thunk):
//   1. bin walk (regions.cu): per-bin control-variate integral `approximation` and region count, regions visited in
//      table order, region patches staged through shared memory — no bin->region CSR is materialised
//      (1.0e9 (bin,region) pairs at BASELINE config 4);
//   2. per-sample region choice: rank ~ U[0, count) from Philox, rank -> region id by a second walk over the tile lists
//      (64-bit occupancy masks per staged chunk);
//   3. residual samples: uniform point in bin ∩ region, weight vol(bin ∩ region), interpolant value approximation_at(x);
//   4. f(x) for all samples (eval thunk), then one thread per bin folds its samples into the reference's online
//      variance/covariance accumulator in double and writes  (sum_f - a*sum_app)/n + a*approximation.
// Compiled with --fmad=false; rule arithmetic through explicit round-to-nearest intrinsics (rules.cuh).
#include "regions.h"
#include <viltrum_b200/device/rules.cuh>
#include <viltrum_b200/device/philox.cuh>
#include <viltrum_b200/device/f32x2.cuh>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cub/device/device_radix_sort.cuh>

using namespace vb200;
namespace R = viltrum::b200::device::rules;
using viltrum::b200::u32x4;
using viltrum::b200::philox4x32;

namespace {

struct TileGeomCv { uint32_t tile[3], tiles[3], res[3]; int db; };
TileGeomCv make_geom_cv(const BinWalk& w, const vb200_domain& dom) {
    TileGeomCv g; g.db = w.db;
    for (int d = 0; d < 3; ++d) { g.tile[d] = w.tile[d]; g.tiles[d] = w.tiles[d]; g.res[d] = d < w.db ? uint32_t(dom.res[d]) : 1u; }
    return g;
}

// rank of every residual sample among its bin's regions: uniform_int_distribution(0, count-1) -> mulhi(u32, count)
// (the reference's Lemire rejection step removes a bias of at most count/2^32, region-russian-roulette.h:14,18-21)
__global__ void cv_ranks_kernel(uint64_t begin, uint64_t nb, uint32_t spp, uint32_t k0, uint32_t k1, const uint32_t* __restrict__ count, uint32_t* __restrict__ rank) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nb * spp) return;
    const uint64_t b = i % nb; const uint32_t j = uint32_t(i / nb);            // sample-major layout: [j][bin]
    const uint64_t bin = begin + b;
    const u32x4 r = philox4x32<10>(u32x4{uint32_t(bin), uint32_t(bin >> 32), j, 0u}, k0, k1);
    rank[i] = __umulhi(r.x, count[b]);
}

// region_stratification_uniform::samples_per_region (reference src/control-variates/region-stratification.h:9-25), the allocation of
// RegionsIntegratorParallelVarianceReductionOptimized (…-variance-reduction-optimized.h:120-133): every region of the bin takes
// spp / n samples and the spp % n left over go to the regions start, start+1, ... (mod n) for ONE random start per bin; regions are
// visited in list order, so sample j belongs to the region whose running sample count first exceeds j.  rank = that region's position
// in the bin's list, rrf = double(spp) / double(samples of that region) — the factor the residual terms carry upstream (:127-128).
__global__ void cv_stratified_ranks_kernel(uint64_t begin, uint64_t nb, uint32_t spp, uint32_t k0, uint32_t k1, const uint32_t* __restrict__ count,
                                           uint32_t* __restrict__ rank, double* __restrict__ rrf) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nb * spp) return;
    const uint64_t b = i % nb; const uint32_t j = uint32_t(i / nb);
    const uint64_t bin = begin + b;
    const uint32_t n = count[b];
    if (n == 0) { rank[i] = 0; rrf[i] = 1.0; return; }
    const uint32_t base = spp / n, rem = spp % n;
    const uint32_t start = rem > 0 ? __umulhi(philox4x32<10>(u32x4{uint32_t(bin), uint32_t(bin >> 32), 0xfffffffeu, 0u}, k0, k1).x, n) : 0u;
    // samples held by the regions before position p: p*base + |[0,p) ∩ ([start, start+rem) mod n)|
    auto before = [&] (uint32_t p) -> uint64_t {
        const uint32_t e = start + rem;                     // < 2n
        uint32_t extra = e <= n ? (p > start ? min(p, e) - start : 0u) : ((p > start ? p - start : 0u) + min(p, e - n));
        return uint64_t(p) * base + extra;
    };
    uint32_t lo = 0, hi = n;                                // largest p with before(p) <= j
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (before(mid) <= j) lo = mid; else hi = mid; }
    const uint32_t mine = uint32_t(before(lo + 1) - before(lo));
    rank[i] = lo;
    rrf[i] = double(spp) / double(mine);
}

// weighted roulettes (rr_integral_region / rr_error_region): the raw 32-bit draw of every residual sample; the walk turns it into
// u * sum(w') and picks the region by inverse CDF (regions.cu walk_rr_kernel)
__global__ void cv_raw_kernel(uint64_t begin, uint64_t nb, uint32_t spp, uint32_t k0, uint32_t k1, uint32_t* __restrict__ raw) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nb * spp) return;
    const uint64_t b = i % nb; const uint32_t j = uint32_t(i / nb);
    const uint64_t bin = begin + b;
    raw[i] = philox4x32<10>(u32x4{uint32_t(bin), uint32_t(bin >> 32), j, 0u}, k0, k1).x;
}

// rank -> region id.  One CTA per bin tile, one thread per bin; the tile's ordered region list is staged in chunks of 64
// pixel boxes, each thread builds the 64-bit mask of the regions that contain its bin and resolves the ranks that fall
// inside the chunk with find-nth-set.
template<int DB>
__global__ void __launch_bounds__(256) cv_resolve_kernel(TileGeomCv g, uint64_t cap, uint64_t begin, uint64_t end, uint32_t spp,
                                                         const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ pend,
                                                         const uint64_t* __restrict__ offsets, const uint32_t* __restrict__ list,
                                                         const uint32_t* __restrict__ rank, uint32_t* __restrict__ chosen) {
    __shared__ uint32_t s_ps[64][DB], s_pe[64][DB], s_id[64];
    const uint64_t t = blockIdx.x, nb = end - begin;
    uint32_t o[3]; { uint64_t q = t; for (int d = 0; d < 3; ++d) { o[d] = uint32_t(q % g.tiles[d]) * g.tile[d]; q /= g.tiles[d]; } }
    uint32_t pos[3] = {0, 0, 0}; { uint32_t k = threadIdx.x; for (int d = 0; d < DB; ++d) { pos[d] = o[d] + k % g.tile[d]; k /= g.tile[d]; } }
    bool live = true; uint64_t bin = 0, prod = 1;
    for (int d = 0; d < DB; ++d) { live = live && pos[d] < g.res[d]; bin += uint64_t(pos[d]) * prod; prod *= g.res[d]; }
    live = live && bin >= begin && bin < end;
    // does any bin of this tile belong to the shard?  (uniform per CTA)
    if (!__syncthreads_or(live ? 1 : 0)) return;
    const uint64_t lo = offsets[t], hi = offsets[t + 1];
    const uint64_t b = bin - begin;
    uint32_t c0 = 0;
    for (uint64_t base = lo; base < hi; base += 64) {
        const int n = int(min(uint64_t(64), hi - base));
        __syncthreads();
        if (int(threadIdx.x) < n) {
            const uint32_t r = list[base + threadIdx.x]; s_id[threadIdx.x] = r;
            for (int d = 0; d < DB; ++d) { s_ps[threadIdx.x][d] = pstart[uint64_t(d) * cap + r]; s_pe[threadIdx.x][d] = pend[uint64_t(d) * cap + r]; }
        }
        __syncthreads();
        if (!live) continue;
        uint32_t m0 = 0, m1 = 0;
        for (int j = 0; j < n; ++j) {
            bool inside = true;
#pragma unroll
            for (int d = 0; d < DB; ++d) inside = inside && pos[d] >= s_ps[j][d] && pos[d] < s_pe[j][d];
            if (inside) { if (j < 32) m0 |= 1u << j; else m1 |= 1u << (j - 32); }
        }
        const uint32_t n0 = __popc(m0), c1 = c0 + n0 + __popc(m1);
        if (c1 > c0) {
            for (uint32_t j = 0; j < spp; ++j) {
                const uint32_t rk = rank[uint64_t(j) * nb + b];
                if (rk >= c0 && rk < c1) {
                    const uint32_t k = rk - c0;
                    const int bit = k < n0 ? __fns(m0, 0, int(k) + 1) : 32 + __fns(m1, 0, int(k - n0) + 1);
                    chosen[uint64_t(j) * nb + b] = s_id[bit];
                }
            }
        }
        c0 = c1;
    }
}

// The same resolution with the masks of GROUP consecutive chunks kept per thread: the tile list is walked once per group to build the
// masks and their running counts, then every sample finds its chunk among GROUP counts and its region inside one mask — instead of
// every chunk looking at every sample (C4: ~16 chunks x 64 samples per bin).  Same chosen[] as cv_resolve_kernel (tested).
template<int DB, int GROUP>
__global__ void __launch_bounds__(256) cv_resolve_grouped_kernel(TileGeomCv g, uint64_t cap, uint64_t begin, uint64_t end, uint32_t spp,
                                                                 const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ pend,
                                                                 const uint64_t* __restrict__ offsets, const uint32_t* __restrict__ list,
                                                                 const uint32_t* __restrict__ rank, uint32_t* __restrict__ chosen) {
    __shared__ uint32_t s_ps[64][DB], s_pe[64][DB];
    const uint64_t t = blockIdx.x, nb = end - begin;
    uint32_t o[3]; { uint64_t q = t; for (int d = 0; d < 3; ++d) { o[d] = uint32_t(q % g.tiles[d]) * g.tile[d]; q /= g.tiles[d]; } }
    uint32_t pos[3] = {0, 0, 0}; { uint32_t k = threadIdx.x; for (int d = 0; d < DB; ++d) { pos[d] = o[d] + k % g.tile[d]; k /= g.tile[d]; } }
    bool live = true; uint64_t bin = 0, prod = 1;
    for (int d = 0; d < DB; ++d) { live = live && pos[d] < g.res[d]; bin += uint64_t(pos[d]) * prod; prod *= g.res[d]; }
    live = live && bin >= begin && bin < end;
    if (!__syncthreads_or(live ? 1 : 0)) return;
    const uint64_t lo = offsets[t], hi = offsets[t + 1];
    const uint64_t b = bin - begin;
    uint32_t c0 = 0;
    for (uint64_t gbase = lo; gbase < hi; gbase += 64ull * GROUP) {
        uint64_t mask[GROUP]; uint32_t cum[GROUP];       // cum[c] = regions containing the bin before chunk c of this group (+ c0)
        uint32_t run = c0;
#pragma unroll
        for (int c = 0; c < GROUP; ++c) {
            const uint64_t base = gbase + 64ull * c;
            const int n = base < hi ? int(min(uint64_t(64), hi - base)) : 0;
            __syncthreads();
            if (int(threadIdx.x) < n) {
                const uint32_t r = list[base + threadIdx.x];
                for (int d = 0; d < DB; ++d) { s_ps[threadIdx.x][d] = pstart[uint64_t(d) * cap + r]; s_pe[threadIdx.x][d] = pend[uint64_t(d) * cap + r]; }
            }
            __syncthreads();
            uint32_t m0 = 0, m1 = 0;
            if (live) {
                for (int j = 0; j < n; ++j) {
                    bool inside = true;
#pragma unroll
                    for (int d = 0; d < DB; ++d) inside = inside && pos[d] >= s_ps[j][d] && pos[d] < s_pe[j][d];
                    if (inside) { if (j < 32) m0 |= 1u << j; else m1 |= 1u << (j - 32); }
                }
            }
            mask[c] = uint64_t(m0) | (uint64_t(m1) << 32);
            cum[c] = run; run += __popc(m0) + __popc(m1);
        }
        if (live && run > c0) {
            for (uint32_t j = 0; j < spp; ++j) {
                const uint32_t rk = rank[uint64_t(j) * nb + b];
                if (rk >= c0 && rk < run) {
                    int c = 0;
#pragma unroll
                    for (int q = 1; q < GROUP; ++q) c += (rk >= cum[q]) ? 1 : 0;      // cum is non-decreasing: the last chunk that starts at or before rk
                    const uint32_t k = rk - cum[c];
                    const uint32_t m0 = uint32_t(mask[c]), m1 = uint32_t(mask[c] >> 32), n0 = __popc(m0);
                    const int bit = k < n0 ? __fns(m0, 0, int(k) + 1) : 32 + __fns(m1, 0, int(k - n0) + 1);
                    chosen[uint64_t(j) * nb + b] = list[gbase + 64ull * c + bit];
                }
            }
        }
        c0 = run;
    }
}

// SoA [sd][cap] -> region-major [n][sd], so that approximation_at reads one region as a few contiguous lines
__global__ void regions_to_aos_kernel(uint64_t n, uint64_t cap, int sd, const float* __restrict__ soa, float* __restrict__ aos) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n * uint64_t(sd)) return;
    const uint64_t r = i / uint64_t(sd); const int k = int(i % uint64_t(sd));
    aos[i] = soa[uint64_t(k) * cap + r];
}

// Region::approximation_at -> app_at (region.h:86-112): fold quadrature.at(t_d, line) over dimension 0, then 1, ... ; float Horner.
// Evaluated depth-first (the value at level d needs S values of level d-1), so only S floats per level are live — registers, no
// local-memory scratch — while every at() sees exactly the inputs the reference's lazy folds give it.
template<int S, int LEVEL> struct ApproxLevel {
    static constexpr int STRIDE = R::ipow(S, LEVEL);       // distance between consecutive indices of dimension LEVEL
    __device__ __forceinline__ static float eval(const float* __restrict__ data, const float* t) {
        float v[S];
#pragma unroll
        for (int e = 0; e < S; ++e) v[e] = ApproxLevel<S, LEVEL - 1>::eval(data + e * STRIDE, t);
        return R::at<S>(t[LEVEL], v);
    }
};
template<int S> struct ApproxLevel<S, 0> {
    __device__ __forceinline__ static float eval(const float* __restrict__ data, const float* t) {
        float v[S];
#pragma unroll
        for (int e = 0; e < S; ++e) v[e] = __ldg(data + e);
        return R::at<S>(t[0], v);
    }
};
template<int S, int D>
__device__ __forceinline__ float approximation_at(const float* __restrict__ data, const float (&t)[D]) {
    return R::fm(ApproxLevel<S, D - 1>::eval(data, t), 1.0f);      // * volume_from(DIM) == 1  (region.h:47-51,91-92)
}

// Throughput form of approximation_at (FAST integrands: statistical parity only): the tensor-product interpolant as a separable
// contraction with the Lagrange basis of the rule's nodes k/(S-1) — S + S^2 + ... + S^D FMAs (363 at S = 3, D = 5: SURVEY.md §8d's
// F_contract) instead of a monomial refit plus Horner per line with every rounding spelled out (~1200 operations).
template<int S> __device__ __forceinline__ void lagrange_basis(float t, float* L) {
#pragma unroll
    for (int k = 0; k < S; ++k) {
        float v = 1.0f;
#pragma unroll
        for (int m = 0; m < S; ++m) if (m != k) v *= (t * float(S - 1) - float(m)) * (1.0f / float(k - m));
        L[k] = v;
    }
}
template<int S, int LEVEL, bool LDG = true> struct FastLevel {      // LDG: region data in global memory (read-only path); else shared memory
    static constexpr int STRIDE = R::ipow(S, LEVEL);
    __device__ __forceinline__ static float eval(const float* __restrict__ data, const float (*L)[S]) {
        float v = 0.0f;
#pragma unroll
        for (int e = 0; e < S; ++e) v = fmaf(L[LEVEL][e], FastLevel<S, LEVEL - 1, LDG>::eval(data + e * STRIDE, L), v);
        return v;
    }
};
template<int S, bool LDG> struct FastLevel<S, 0, LDG> {
    __device__ __forceinline__ static float eval(const float* __restrict__ data, const float (*L)[S]) {
        float v = 0.0f;
#pragma unroll
        for (int e = 0; e < S; ++e) v = fmaf(L[0][e], LDG ? __ldg(data + e) : data[e], v);
        return v;
    }
};

// ---- importance sampling of a region's interpolant -----------------------------------------------------------
//   Simpson::pdf_points / pdf_unnormalized / cdf / inv_cdf / sample            reference src/newton-cotes/rules.h:104-247
//   Region::sample_subrange / sample_marginal / pdf_subrange / pdf_at           reference src/newton-cotes/region.h:220-343
//   region_sampling_importance / _mis / _russian_roulette                       reference src/control-variates/region-sampling.h:22-135
// A sample is drawn dimension by dimension from the |interpolant| restricted to bin ∩ region: the marginal of dimension k (the other
// dimensions folded with pdf_integral_subrange) is a parabola whose CDF — a cubic — is inverted in closed form (Cardano / Viète), then the
// array is conditioned on the drawn coordinate (fold with pdf_unnormalized) and the next dimension follows.  pow/acos/cos make a bit
// contract impossible: plain fp32 with the library's fast intrinsics, statistical parity only.
namespace imp {
__device__ __forceinline__ void coeff3(const float* p, float* c) { c[0] = p[0]; c[1] = -3.0f * p[0] + 4.0f * p[1] - p[2]; c[2] = 2.0f * p[0] - 4.0f * p[1] + 2.0f * p[2]; }
// rules.h:104-147 with NormDefault: |p|, shifted down by the (negative, interior) minimum of its parabola
__device__ __forceinline__ void pdf_points(const float* p, float* q) {
    q[0] = fabsf(p[0]); q[1] = fabsf(p[1]); q[2] = fabsf(p[2]);
    float c[3]; coeff3(q, c);
    float ymin = 0.0f;
    if (c[2] > 0.0f) {
        const float tmin = -c[1] / (2.0f * c[2]);
        if (tmin > 0.0f && tmin < 1.0f) { const float y = (c[2] * tmin + c[1]) * tmin + c[0]; if (y < ymin) ymin = y; }
    }
    q[0] -= ymin; q[1] -= ymin; q[2] -= ymin;
}
__device__ __forceinline__ float cdf3(float t, const float* c) { return ((c[2] * t * (1.0f / 3.0f) + c[1] * 0.5f) * t + c[0]) * t; }       // rules.h:173-177
__device__ __forceinline__ float pdf_integral(float t0, float t1, const float* p) { float q[3], c[3]; pdf_points(p, q); coeff3(q, c); return cdf3(t1, c) - cdf3(t0, c); }   // :150-154
__device__ __forceinline__ float pdf_unnormalized(float t, const float* p) { float q[3], c[3]; pdf_points(p, q); coeff3(q, c); return (c[2] * t + c[1]) * t + c[0]; }       // :157-161
__device__ __forceinline__ float cuberoot(float x) { return x < 0.0f ? -powf(-x, 1.0f / 3.0f) : powf(x, 1.0f / 3.0f); }
// rules.h:184-247: x = s*cdf(t1) + (1-s)*cdf(t0), roots of cdf(t) - x, the first one inside [t0,t1] (1e-5 slack), else uniform
__device__ float sample1(float s, float t0, float t1, const float* p) {
    float q[3], cf[3]; pdf_points(p, q); coeff3(q, cf);
    const float x = s * cdf3(t1, cf) + (1.0f - s) * cdf3(t0, cf);
    float a = cf[2] * (1.0f / 3.0f), b = cf[1] * 0.5f, c = cf[0], d = -x;
    const float sum = a + b + c + d; a /= sum; b /= sum; c /= sum; d /= sum;
    float sol[3]; int ns = 0;
    if (fabsf(a) < 1.e-3f) {
        if (fabsf(b) < 1.e-3f) { if (fabsf(c) >= 1.e-3f) sol[ns++] = -d / c; }
        else { const float disc = c * c - 4.0f * b * d; if (disc >= 0.0f) { const float sq = sqrtf(disc); sol[ns++] = (-c + sq) / (2.0f * b); sol[ns++] = (-c - sq) / (2.0f * b); } }
    } else {
        const float pp = c / a - b * b / (3.0f * a * a);
        const float qq = 2.0f * b * b * b / (27.0f * a * a * a) - b * c / (3.0f * a * a) + d / a;
        const float sqr = 0.25f * qq * qq + pp * pp * pp * (1.0f / 27.0f);
        if (sqr >= 0.0f) sol[ns++] = cuberoot(-0.5f * qq - sqrtf(sqr)) + cuberoot(-0.5f * qq + sqrtf(sqr)) - b / (3.0f * a);
        else for (int i = 0; i < 3; ++i) sol[ns++] = 2.0f * sqrtf(-pp / 3.0f) * cosf(acosf(3.0f * qq / (2.0f * pp) * sqrtf(-3.0f / pp)) * (1.0f / 3.0f) - 6.2831853071795865f * float(i) / 3.0f) - b / (3.0f * a);
    }
    for (int i = 0; i < ns; ++i) {
        float r = sol[i];
        if (r < t0 && fabsf(t0 - r) < 1.e-5f) r = t0;
        if (r > t1 && fabsf(r - t1) < 1.e-5f) r = t1;
        if (r >= t0 && r <= t1 && !isnan(r)) return r;
    }
    return s * (t1 - t0) + t0;
}
// one region: data[3^D] (dimension 0 fastest), normalised box [a,b]^D of bin ∩ region.  u[D] uniforms -> normalised position pos[D];
// returns pdf_at(pos) / pdf_sub(a,b) = the density of pos in NORMALISED coordinates times 1 (the caller divides by the region's volume).
template<int D>
__device__ float sample_and_pdf(const float* __restrict__ data, const float* a, const float* b, const float* u, bool importance_pos, float* pos, float* pdf_sub_out) {
    constexpr int N = R::ipow(3, D);
    float arr[N];      // the array still to be sampled: dimensions k..D-1, conditioned on pos[0..k)
    for (int i = 0; i < N; ++i) arr[i] = __ldg(data + i);
    float pdf_at = 0.0f, pdf_sub = 0.0f;
    int n = N;
    for (int k = 0; k < D; ++k) {
        // marginal of dimension k: fold the last dimensions with pdf_integral over their [a,b] (region.h:221-240)
        float m[N];
        for (int i = 0; i < n; ++i) m[i] = arr[i];
        int len = n;
        for (int dd = D - 1; dd > k; --dd) {
            const int outn = len / 3;
            for (int o = 0; o < outn; ++o) { const float line[3] = {m[o], m[o + outn], m[o + 2 * outn]}; m[o] = pdf_integral(a[dd], b[dd], line); }
            len = outn;
        }
        // m[0..3) is the marginal along dimension k
        if (k == 0) pdf_sub = pdf_integral(a[0], b[0], m);                                   // region.h:268-279 (pdf_sub: every dimension folded over [a,b])
        pos[k] = importance_pos ? sample1(u[k], a[k], b[k], m) : u[k] * (b[k] - a[k]) + a[k];
        // condition on pos[k]: fold dimension 0 of arr with pdf_unnormalized (region.h:250-254)
        const int outn = n / 3;
        for (int o = 0; o < outn; ++o) { const float line[3] = {arr[3 * o], arr[3 * o + 1], arr[3 * o + 2]}; arr[o] = pdf_unnormalized(pos[k], line); }
        n = outn;
    }
    // pdf_at (region.h:300-313): the folds are not linear (|.| and the shift of pdf_points), so the density is evaluated the reference's way —
    // pdf_unnormalized over the LAST dimension first — rather than read off the conditioning chain above
    for (int i = 0; i < N; ++i) arr[i] = __ldg(data + i);
    int len = N;
    for (int dd = D - 1; dd >= 0; --dd) {
        const int outn = len / 3;
        for (int o = 0; o < outn; ++o) { const float line[3] = {arr[o], arr[o + outn], arr[o + 2 * outn]}; arr[o] = pdf_unnormalized(pos[dd], line); }
        len = outn;
    }
    pdf_at = arr[0];
    *pdf_sub_out = pdf_sub;
    return pdf_at;
}
}

// residual samples under region_sampling_importance (RS = 1), region_sampling_mis (2) or region_sampling_russian_roulette (3); Simpson
// tables.  Same inputs and outputs as cv_samples_kernel; weight = the policy's sample weight (1/pdf, the constant MIS weight, ...).
template<int D, int RS>
__global__ void __launch_bounds__(128) cv_samples_importance_kernel(vb200_domain dom, uint64_t begin, uint64_t nb, uint32_t spp, uint32_t k0, uint32_t k1,
                                                                    uint64_t cap, const float* __restrict__ rmin, const float* __restrict__ rmax, const float* __restrict__ aos,
                                                                    const uint32_t* __restrict__ sorted_region, const uint32_t* __restrict__ sorted_index,
                                                                    float* __restrict__ points, float* __restrict__ weight, float* __restrict__ app, double power, double cutoff) {
    const uint64_t tpos = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t N = nb * spp;
    if (tpos >= N) return;
    const uint64_t i = sorted_index[tpos];
    const uint64_t b = i % nb; const uint32_t j = uint32_t(i / nb);
    const uint64_t bin = begin + b;
    const uint32_t r = sorted_region[tpos];
    uint32_t pos[VB200_MAX_DIMBINS]; { uint64_t q = bin; for (int d = 0; d < dom.dimbins; ++d) { pos[d] = uint32_t(q % dom.res[d]); q /= dom.res[d]; } }
    float lo[D], hi[D], ia[D], ib[D], na[D], nbn[D], u[D + 1], L[D][3];
    float vol = 1.0f, rvol = 1.0f;
    u32x4 rnd{0, 0, 0, 0};
#pragma unroll
    for (int d = 0; d <= D; ++d) {
        if ((d & 3) == 0) rnd = philox4x32<10>(u32x4{uint32_t(bin), uint32_t(bin >> 32), j, uint32_t(1 + d / 4)}, k0, k1);
        u[d] = viltrum::b200::u01((d & 3) == 0 ? rnd.x : (d & 3) == 1 ? rnd.y : (d & 3) == 2 ? rnd.z : rnd.w);
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
        lo[d] = rmin[uint64_t(d) * cap + r]; hi[d] = rmax[uint64_t(d) * cap + r];
        float ba = dom.rmin[d], bb = dom.rmax[d];
        if (d < dom.dimbins) { ba = fmaf(float(pos[d]), dom.drange[d], dom.rmin[d]); bb = fmaf(float(pos[d] + 1u), dom.drange[d], dom.rmin[d]); }
        ia[d] = fmaxf(ba, lo[d]); ib[d] = fmaxf(ia[d], fminf(bb, hi[d]));
        vol *= ib[d] - ia[d]; rvol *= hi[d] - lo[d];
        const float inv = hi[d] > lo[d] ? 1.0f / (hi[d] - lo[d]) : 0.0f;
        na[d] = (ia[d] - lo[d]) * inv; nbn[d] = (ib[d] - lo[d]) * inv;
    }
    const float* data = aos + uint64_t(r) * uint64_t(R::ipow(3, D));
    float x[D], w;
    bool uniform_pos = false;
    if (RS == 1 && vol < 1.e-5f) { uniform_pos = true; w = vol; }                             // region-sampling.h:31-32: tiny subranges are sampled uniformly
    else {
        float t[D], pdf_sub;
        const float pdf_at = imp::sample_and_pdf<D>(data, na, nbn, u, true, t, &pdf_sub);
        const float pdf_imp = pdf_at / (rvol * pdf_sub);                                       // Region::pdf_subrange (region.h:333-337)
        // a region whose interpolant vanishes over the box has no density to sample (0/0: the reference returns NaN there): uniform sampling instead
        if (!(pdf_sub > 0.0f) || !(pdf_imp == pdf_imp) || isinf(pdf_imp)) { uniform_pos = true; w = vol; }
        else if (RS == 1) w = pdf_imp < 1.e-5f ? 0.0f : 1.0f / pdf_imp;                        // :42
        else if (RS == 3) {                                                                    // region_sampling_russian_roulette :46-83
            const float pdf_importance = rvol * pdf_sub, pdf_uniform = vol;
            if (u[D] < pdf_importance / (pdf_importance + pdf_uniform)) w = pdf_imp < 1.e-10f ? 0.0f : 1.0f / pdf_imp;
            else { uniform_pos = true; w = vol; }                                              // pdf = 1/volume
        } else {                                                                               // region_sampling_mis :85-135
            const float pdf_uni = 1.0f / vol;
            float mi = powf(pdf_imp, float(power)), mu = powf(pdf_uni, float(power)), ms = mi + mu;
            if (mi < float(cutoff) * ms) { mi = 0.0f; ms = mu; }
            if (mu < float(cutoff) * ms) { mu = 0.0f; ms = mi; }
            const float p_imp = pdf_imp > 1.e-10f ? mi / pdf_imp : 0.0f, p_uni = mu * vol, p_sum = p_imp + p_uni;
            w = p_sum / ms;
            if (u[D] < p_uni / p_sum) uniform_pos = true;
        }
        if (!uniform_pos) {
#pragma unroll
            for (int d = 0; d < D; ++d) x[d] = fmaf(t[d], hi[d] - lo[d], lo[d]);               // Range::pos_from_range
        }
    }
    if (uniform_pos) {
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = fmaf(u[d], ib[d] - ia[d], ia[d]);
    }
#pragma unroll
    for (int d = 0; d < D; ++d) { points[uint64_t(d) * N + tpos] = x[d]; lagrange_basis<3>(hi[d] > lo[d] ? (x[d] - lo[d]) / (hi[d] - lo[d]) : 0.0f, L[d]); }
    weight[tpos] = w;
    app[tpos] = FastLevel<3, D - 1>::eval(data, L);
}

// residual samples: chosen region -> bin ∩ region box -> uniform point, weight, interpolant value.
// REPLAY: points are given (AoS [bin][spp][D]), only weights/interpolant are computed.
// The samples are visited in REGION-SORTED order (thread t handles sample sorted_index[t] of region sorted_region[t]): the lanes of a
// warp then read the same region — its box and its S^D interpolation samples (972 B at C4) come through L1 as broadcasts instead of
// 32 different gathers per load instruction.  What a sample computes depends only on (bin, sample number, region), so the order
// changes no bit; outputs are written at the sorted position t (coalesced) and brought back by cv_unsort_kernel.
template<int S, int D, bool REPLAY, bool FAST = false>
__global__ void __launch_bounds__(128) cv_samples_kernel(vb200_domain dom, uint64_t begin, uint64_t nb, uint32_t spp, uint32_t k0, uint32_t k1,
                                                         uint64_t cap, const float* __restrict__ rmin, const float* __restrict__ rmax, const float* __restrict__ aos,
                                                         const uint32_t* __restrict__ sorted_region, const uint32_t* __restrict__ sorted_index, const float* __restrict__ replay_points,
                                                         float* __restrict__ points, float* __restrict__ weight, float* __restrict__ app) {
    const uint64_t tpos = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t N = nb * spp;
    if (tpos >= N) return;
    const uint64_t i = sorted_index[tpos];
    const uint64_t b = i % nb; const uint32_t j = uint32_t(i / nb);
    const uint64_t bin = begin + b;
    const uint32_t r = sorted_region[tpos];
    // bin box in the binned dims (…-variance-reduction.h:71-73), region extent elsewhere; Range::intersection (range.h:92-101)
    uint32_t pos[VB200_MAX_DIMBINS]; { uint64_t q = bin; for (int d = 0; d < dom.dimbins; ++d) { pos[d] = uint32_t(q % dom.res[d]); q /= dom.res[d]; } }
    float a[D], w[D], t[D], x[D];
    float vol = 1.0f;
    u32x4 rnd{0, 0, 0, 0};
    if constexpr (FAST) {        // plain fp32 + FMA; the same samples (same Philox words), the same estimator
        float L[D][S];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float lo = rmin[uint64_t(d) * cap + r], hi = rmax[uint64_t(d) * cap + r];
            float ba = dom.rmin[d], bb = dom.rmax[d];
            if (d < dom.dimbins) { ba = fmaf(float(pos[d]), dom.drange[d], dom.rmin[d]); bb = fmaf(float(pos[d] + 1u), dom.drange[d], dom.rmin[d]); }
            const float ia = fmaxf(ba, lo), ib = fmaxf(ia, fminf(bb, hi));
            const float wd = ib - ia;
            vol *= wd;
            if ((d & 3) == 0) rnd = philox4x32<10>(u32x4{uint32_t(bin), uint32_t(bin >> 32), j, uint32_t(1 + d / 4)}, k0, k1);
            const uint32_t u = (d & 3) == 0 ? rnd.x : (d & 3) == 1 ? rnd.y : (d & 3) == 2 ? rnd.z : rnd.w;
            const float xd = fmaf(viltrum::b200::u01(u), wd, ia);
            points[uint64_t(d) * N + tpos] = xd;
            lagrange_basis<S>(hi > lo ? (xd - lo) / (hi - lo) : 0.0f, L[d]);
        }
        weight[tpos] = vol;
        app[tpos] = FastLevel<S, D - 1>::eval(aos + uint64_t(r) * uint64_t(R::ipow(S, D)), L);
        return;
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const float lo = rmin[uint64_t(d) * cap + r], hi = rmax[uint64_t(d) * cap + r];
        float ba = dom.rmin[d], bb = dom.rmax[d];
        if (d < dom.dimbins) { ba = R::fa(dom.rmin[d], R::fm(float(pos[d]), dom.drange[d])); bb = R::fa(dom.rmin[d], R::fm(float(pos[d] + 1u), dom.drange[d])); }
        const float ia = fmaxf(ba, lo), ib = fmaxf(ia, fminf(bb, hi));
        a[d] = ia; w[d] = R::fs(ib, ia);
        vol = R::fm(vol, w[d]);                                                   // Range::volume of bin ∩ region (region-sampling.h:18)
        if (REPLAY) x[d] = replay_points[(b * spp + j) * D + d];
        else {
            if ((d & 3) == 0) rnd = philox4x32<10>(u32x4{uint32_t(bin), uint32_t(bin >> 32), j, uint32_t(1 + d / 4)}, k0, k1);
            const uint32_t u = (d & 3) == 0 ? rnd.x : (d & 3) == 1 ? rnd.y : (d & 3) == 2 ? rnd.z : rnd.w;
            x[d] = R::fa(R::fm(viltrum::b200::u01(u), w[d]), ia);                 // uniform_real_distribution: u*(b-a)+a (region-sampling.h:13-17)
        }
        t[d] = R::pos_in_range(lo, hi, x[d]);
        points[uint64_t(d) * N + tpos] = x[d];
    }
    weight[tpos] = vol;
    app[tpos] = approximation_at<S, D>(aos + uint64_t(r) * uint64_t(R::ipow(S, D)), t);
}

__global__ void cv_iota_kernel(uint64_t n, uint32_t* __restrict__ idx) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = uint32_t(i);
}

// region-sorted position -> sample-major position [j][bin]: one 16-byte record (f, interpolant, weight) per sample
__global__ void cv_unsort_kernel(uint64_t n, const uint32_t* __restrict__ sorted_index, const float* __restrict__ fval, const float* __restrict__ app,
                                 const float* __restrict__ weight, float4* __restrict__ rec) {
    const uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t < n) rec[sorted_index[t]] = make_float4(fval[t], app[t], weight[t], 0.0f);
}

// cv_optimize_weight::Accumulator (weight-strategy.h:40-110) / cv_fixed_weight::Accumulator (:7-35): samples in order, moments in double
struct CvAccumulator {
    float sum_f = 0.0f, sum_app = 0.0f, fixed_sum = 0.0f; uint64_t size = 0;
    double k_f = 0, k_app = 0, e_f = 0, e_ap = 0, e_ap2 = 0, e_fap = 0;
    // f(sample)*double(factor)*rrfactor*sfactor, rounded to the Sample type (…-variance-reduction.h:97-100)
    __device__ __forceinline__ void push(float f, float app, float sfactor, double factor, double rrfactor, double fixed_alpha) {
        const double sf = double(sfactor);
        const float fs = R::d2f(R::dm(R::dm(R::dm(double(f), factor), rrfactor), sf));
        const float as = R::d2f(R::dm(R::dm(R::dm(double(app), factor), rrfactor), sf));
        const double nf = double(fabsf(fs)), na = double(fabsf(as));                     // NormDefault (norm.h:12)
        if (size == 0) { k_f = nf; k_app = na; }
        e_f = R::da(e_f, R::ds(nf, k_f));
        e_ap = R::da(e_ap, R::ds(na, k_app));
        e_ap2 = R::da(e_ap2, R::dm(R::ds(na, k_app), R::ds(na, k_app)));
        e_fap = R::da(e_fap, R::dm(R::ds(nf, k_f), R::ds(na, k_app)));
        sum_f = R::fa(sum_f, fs); sum_app = R::fa(sum_app, as);
        fixed_sum = R::d2f(R::da(double(fixed_sum), R::ds(double(fs), R::dm(fixed_alpha, double(as)))));      // sum += f - alpha*app (:20)
        ++size;
    }
    __device__ __forceinline__ float result(float approximation, int fixed_weight, double fixed_alpha) const {
        if (fixed_weight) return size == 0 ? approximation : R::d2f(R::da(R::dd(double(fixed_sum), double(size)), R::dm(fixed_alpha, double(approximation))));   // :24-27
        if (size < 2) return approximation;                                              // weight-strategy.h:95
        const double n = double(size), n1 = double(size - 1);
        const double covariance = R::dd(R::ds(e_fap, R::dd(R::dm(e_f, e_ap), n)), n1);
        const double variance = R::dd(R::ds(e_ap2, R::dd(R::dm(e_ap, e_ap), n)), n1);
        double alpha;
        const double v = fmax(0.0, variance);
        if (v <= 0.0) alpha = 1.0;
        else alpha = R::dd(fmin(v, fmax(0.0, covariance)), v);
        return R::d2f(R::da(R::dd(R::ds(double(sum_f), R::dm(alpha, double(sum_app))), n), R::dm(alpha, double(approximation))));
    }
};

// one thread per bin folds its samples (sample-major records) in order
__global__ void __launch_bounds__(128) cv_accumulate_kernel(uint64_t begin, uint64_t nb, uint32_t spp, uint64_t nbins_total,
                                                            const uint32_t* __restrict__ count, const float* __restrict__ approx,
                                                            const float4* __restrict__ rec /* (f, interpolant, weight) */,
                                                            float* __restrict__ out, int fixed_weight, double fixed_alpha, const double* __restrict__ rrf) {
    const uint64_t b = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const double factor = double(nbins_total), rr_uniform = double(count[b]);         // rr_uniform_region: rr.max()+1 (region-russian-roulette.h:19)
    CvAccumulator acc;
    for (uint32_t j = 0; j < spp; ++j) {
        const uint64_t i = uint64_t(j) * nb + b;
        const float4 smp = rec[i];
        acc.push(smp.x, smp.y, smp.z, factor, rrf ? rrf[i] : rr_uniform, fixed_alpha);  // weighted roulettes: 1/probability of the chosen region
    }
    out[begin + b] = acc.result(approx[b], fixed_weight, fixed_alpha);                   // '=' (…-variance-reduction.h:102)
}

// The same contraction over a region record in SHARED memory, read as 128-bit words: the S^D values are consumed strictly in storage order
// (dimension 0 fastest), every level keeps one running sum.  Same operations in the same order as FastLevel (identical bits), a quarter
// of the shared-memory wavefronts — which is what bounds the tile-major residual kernel.  `data` must be 16-byte aligned and readable up to
// the next multiple of four values.
template<int S, int D>
__device__ __forceinline__ float fast_eval_stream(const float* __restrict__ data, const float (*L)[S]) {
    constexpr int SD = R::ipow(S, D);
    const float4* d4 = reinterpret_cast<const float4*>(data);
    float acc[D];
#pragma unroll
    for (int l = 0; l < D; ++l) acc[l] = 0.0f;
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < SD; ++i) {
        if ((i & 3) == 0) c = d4[i >> 2];
        const float x = (i & 3) == 0 ? c.x : (i & 3) == 1 ? c.y : (i & 3) == 2 ? c.z : c.w;
        acc[0] = fmaf(L[0][i % S], x, acc[0]);
        int q = i;
#pragma unroll
        for (int l = 1; l < D; ++l) {
            if (q % S != S - 1) break;
            q /= S;
            acc[l] = fmaf(L[l][q % S], acc[l - 1], acc[l]);
            acc[l - 1] = 0.0f;
        }
    }
    return acc[D - 1];
}

// Layout of a region record in the residual kernel's shared-memory slots.  The S slabs of the slowest dimension are stored in PAIRS — slab 2m and
// slab 2m+1 interleaved value by value — so that the contraction of a pair of slabs runs on packed FP32 (FFMA2: one instruction, both slabs), an odd
// last slab stays scalar: S = 3, D = 5 takes 120 FFMA2 + 123 FFMA instead of 363 FFMA.
template<int S, int D> struct CvSlot {
    static constexpr int SD = R::ipow(S, D), H = R::ipow(S, D - 1), NPAIR = S / 2, LO = S % 2;
    static constexpr int PS = (2 * H + 3) / 4 * 4;                 // words of a pair of slabs (16-byte aligned)
    static constexpr int LOFF = NPAIR * PS;                         // the unpaired last slab
    static constexpr int DW = (LOFF + LO * H + 3) / 4 * 4;         // data words; the ranges follow
    __host__ __device__ static constexpr int pos(int i) {           // sample i (dimension 0 fastest) -> word of the slot
        return (i / H) < 2 * NPAIR ? ((i / H) / 2) * PS + 2 * (i % H) + ((i / H) & 1) : LOFF + (i % H);
    }
};
template<int S, int D>
__device__ __forceinline__ float fast_eval_pairs(const float* __restrict__ reg, const float (*L)[S]) {
    using SL = CvSlot<S, D>;
    using viltrum::b200::f32x2;
    float v = 0.0f;
    f32x2 L2[D > 1 ? D - 1 : 1][S];
#pragma unroll
    for (int l = 0; l < D - 1; ++l)
#pragma unroll
        for (int e = 0; e < S; ++e) L2[l][e] = f32x2(L[l][e]);
#pragma unroll
    for (int m = 0; m < SL::NPAIR; ++m) {
        const ulonglong2* d2 = reinterpret_cast<const ulonglong2*>(reg + m * SL::PS);
        f32x2 acc[D > 1 ? D - 1 : 1];
#pragma unroll
        for (int l = 0; l < D - 1; ++l) acc[l] = f32x2(0.0f);
        ulonglong2 c = make_ulonglong2(0ull, 0ull);
#pragma unroll
        for (int p = 0; p < SL::H; ++p) {
            if ((p & 1) == 0) c = d2[p >> 1];
            f32x2 x; x.v = (p & 1) ? c.y : c.x;                   // (slab 2m, slab 2m+1) at position p
            acc[0] = viltrum::b200::mad(L2[0][p % S], x, acc[0]);
            int q = p;
#pragma unroll
            for (int l = 1; l < D - 1; ++l) {
                if (q % S != S - 1) break;
                q /= S;
                acc[l] = viltrum::b200::mad(L2[l][q % S], acc[l - 1], acc[l]);
                acc[l - 1] = f32x2(0.0f);
            }
        }
        v = fmaf(L[D - 1][2 * m], acc[D - 2].lo(), v);
        v = fmaf(L[D - 1][2 * m + 1], acc[D - 2].hi(), v);
    }
    if constexpr (SL::LO == 1) v = fmaf(L[D - 1][S - 1], fast_eval_stream<S, D - 1>(reg + SL::LOFF, L), v);
    return v;
}

// region-major copy of the ranges, ranges[r][0..D) = min, [D..2D) = max: the residual kernel stages a region's box with one 8*D-byte read
__global__ void ranges_to_aos_kernel(uint64_t n, uint64_t cap, int dim, const float* __restrict__ rmin, const float* __restrict__ rmax, float* __restrict__ out) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n * uint64_t(2 * dim)) return;
    const uint64_t r = i / uint64_t(2 * dim); const int k = int(i % uint64_t(2 * dim));
    out[i] = k < dim ? rmin[uint64_t(k) * cap + r] : rmax[uint64_t(k - dim) * cap + r];
}

// ---- tile-major residual pass (throughput path: FAST integrand, rr_uniform_region, 2-D bin grid) ------------------------------------
// The sample-major pipeline above sorts ALL residual samples of a slab by region with a device-wide radix sort, evaluates them in that
// order and scatters 16-byte records back (sort 1.5 ms + un-sort 2.7 ms of BASELINE config 4's 23 ms).  Here a CTA owns one
// 16x16 bin tile and does, per pass of J samples per bin and entirely in shared memory: the draw of every sample's region
// from the tile's region list, a counting sort of the pass's samples by their position in that list (so that consecutive
// threads evaluate the same region: its box and interpolation samples arrive as L1 broadcasts), and the sample points / weights /
// interpolant values, which go to global memory in that tile-local order together with the sample each slot belongs to.  After the
// integrand's launch over those points, cv_tile_accumulate_kernel brings a tile's values back into sample order through shared memory
// and one thread per bin folds them in the reference's order.  No device-wide sort, no scattered global traffic.
constexpr int CVT_BINS = 256, CVT_MAXLIST = 4096, CVT_MAXPASS = 64, CVT_ACCPASS = 32, CVT_SLOTS = 4;
constexpr int cvt_slot_words(int sd, int d) { return ((sd + 2 * d + 7) / 8) * 8 + 4; }
struct CvTileArgs {
    TileGeomCv g; vb200_domain dom; uint64_t cap, begin, end, tile0, nbins_total; uint32_t spp, J, k0, k1;
    const uint32_t* pstart; const uint32_t* pend; const uint64_t* offsets; const uint32_t* list; const uint32_t* count;      // count: indexed by bin - count_base
    uint64_t count_base;
    const float* ranges; const float* aos;            // region-major: ranges[r][2*D], aos[r][S^D]
    float* points; float* weight; float* app; unsigned short* owner;      // [slots] per array, points SoA [d][slots]; slots = tiles * 256 * spp
    uint64_t slots;
};
#ifndef VB200_CVT_MINB
#define VB200_CVT_MINB 2      // resident CTAs per SM the residual kernel is compiled for (shared memory allows two; measured 1: 2.98, 2: 2.10, 3: 2.12 ms per launch, profiles/results_r2.md)
#endif
template<int S, int D>
__global__ void __launch_bounds__(256, VB200_CVT_MINB) cv_tile_samples_kernel(const CvTileArgs a) {
    extern __shared__ __align__(16) unsigned char cvt_smem[];
    unsigned short* s_rank   = reinterpret_cast<unsigned short*>(cvt_smem);                 // [J][256] the pass's sample ids sorted by list position
    unsigned short* s_choice = s_rank + a.J * CVT_BINS;                              // [J][256] position in the tile's region list (0xffff = no sample)
    unsigned short* s_hist   = s_choice + a.J * CVT_BINS;                            // [L] samples per list entry, then their running offsets
    uint32_t* s_box = reinterpret_cast<uint32_t*>(s_hist + CVT_MAXLIST);                     // [L] pixel boxes of the list entries, tile-local bytes
    float* s_region = reinterpret_cast<float*>(s_box);                                       // [8 warps][CVT_SLOTS] region records of the evaluation phase (the boxes are done with by then)
    __shared__ uint32_t s_scan[CVT_BINS];
    const uint64_t t = a.tile0 + blockIdx.x;
    const TileGeomCv& g = a.g;
    const uint32_t tid = threadIdx.x;
    uint32_t o[2]; o[0] = uint32_t(t % g.tiles[0]) * g.tile[0]; o[1] = uint32_t(t / g.tiles[0]) * g.tile[1];
    uint32_t pos[2]; pos[0] = o[0] + tid % g.tile[0]; pos[1] = o[1] + tid / g.tile[0];
    const uint64_t bin = uint64_t(pos[0]) + uint64_t(pos[1]) * g.res[0];
    const bool live = pos[0] < g.res[0] && pos[1] < g.res[1] && bin >= a.begin && bin < a.end;
    const uint32_t cnt = live ? a.count[bin - a.count_base] : 0u;
    const uint64_t lo = a.offsets[t], hi = a.offsets[t + 1];
    const uint32_t L = uint32_t(hi - lo);
    const uint64_t slot_tile = uint64_t(blockIdx.x) * CVT_BINS * a.spp;
    for (uint32_t j0 = 0; j0 < a.spp; j0 += a.J) {
        const uint32_t J = min(a.J, a.spp - j0);
        // 1.+2. region of every sample: rr_uniform_region picks uniformly among the regions that touch the bin (region-russian-roulette.h:9-28).
        //    Drawn here by rejection from the TILE's list — a uniform candidate entry is accepted iff its pixel box contains the bin — which
        //    is uniform over the bin's own regions without ever enumerating them (~80 % of a tile's entries contain a given bin at BASELINE
        //    config 4: 1.25 candidates per sample).  A sample still without a region after 17 candidates
        //    (a bin that few of its tile's regions touch) falls back to rank + linear scan, exactly as the sample-major pipeline resolves it.
        for (uint32_t i = tid; i < L; i += CVT_BINS) {
            const uint32_t r = a.list[lo + i];
            const uint32_t x0 = max(a.pstart[r], o[0]) - o[0], x1 = min(a.pend[r], o[0] + 16u) - o[0];
            const uint32_t y0 = max(a.pstart[a.cap + r], o[1]) - o[1], y1 = min(a.pend[a.cap + r], o[1] + 16u) - o[1];
            s_box[i] = x0 | (x1 << 8) | (y0 << 16) | (y1 << 24);                // the entry's pixel box inside the tile, one byte per edge
            s_hist[i] = 0;
        }
        __syncthreads();
        const uint32_t bx = tid % 16u, by = tid / 16u;
        auto contains = [&] (uint32_t i) -> bool {
            const uint32_t bxw = s_box[i];
            return bx >= (bxw & 255u) && bx < ((bxw >> 8) & 255u) && by >= ((bxw >> 16) & 255u) && by < (bxw >> 24);
        };
        u32x4 first{0, 0, 0, 0};
        for (uint32_t j = 0; j < J; ++j) {
            uint32_t idx = 0xffffu;
            if (cnt > 0u) {
                // the first candidate of four consecutive samples comes from ONE Philox block (most samples accept it); later candidates from the sample's own blocks
                const uint32_t gj = j0 + j;                   // sample number within the bin: word gj % 4 of block gj / 4
                if ((gj & 3u) == 0u || j == 0u) first = philox4x32<10>(u32x4{uint32_t(bin), uint32_t(bin >> 32), gj >> 2, 0x7fu}, a.k0, a.k1);
                {
                    const uint32_t w = (gj & 3u) == 0u ? first.x : (gj & 3u) == 1u ? first.y : (gj & 3u) == 2u ? first.z : first.w;
                    const uint32_t cand0 = __umulhi(w, L);
                    if (contains(cand0)) idx = cand0;
                }
                for (uint32_t blk = 0; blk < 4u && idx == 0xffffu; ++blk) {
                    const u32x4 c = philox4x32<10>(u32x4{uint32_t(bin), uint32_t(bin >> 32), j0 + j, 0x80u + blk}, a.k0, a.k1);
                    const uint32_t cand[4] = {__umulhi(c.x, L), __umulhi(c.y, L), __umulhi(c.z, L), __umulhi(c.w, L)};
#pragma unroll
                    for (int q = 0; q < 4; ++q) if (idx == 0xffffu && contains(cand[q])) idx = cand[q];
                }
                if (idx == 0xffffu) {          // rank among the bin's regions, resolved by a scan of the list (rare)
                    uint32_t rk = __umulhi(philox4x32<10>(u32x4{uint32_t(bin), uint32_t(bin >> 32), j0 + j, 0u}, a.k0, a.k1).x, cnt);
                    for (uint32_t i = 0; i < L; ++i) if (contains(i)) { if (rk == 0u) { idx = i; break; } --rk; }
                }
            }
            s_choice[j * CVT_BINS + tid] = (unsigned short)idx;
            if (idx != 0xffffu) atomicAdd(reinterpret_cast<unsigned int*>(s_hist) + (idx >> 1), (idx & 1u) ? 0x10000u : 1u);      // 16-bit counters, two per word
        }
        __syncthreads();
        // 3. counting sort of the pass's samples by list position: exclusive scan of the counters, then every sample takes its slot
        {
            const uint32_t per = (L + CVT_BINS - 1) / CVT_BINS, b0 = tid * per, b1 = min(b0 + per, L);
            uint32_t sum = 0; for (uint32_t i = b0; i < b1; ++i) sum += s_hist[i];
            s_scan[tid] = sum; __syncthreads();
            for (uint32_t off = 1; off < CVT_BINS; off <<= 1) { const uint32_t v = tid >= off ? s_scan[tid - off] : 0u; __syncthreads(); s_scan[tid] += v; __syncthreads(); }
            uint32_t run = s_scan[tid] - sum;
            for (uint32_t i = b0; i < b1; ++i) { const uint32_t c = s_hist[i]; s_hist[i] = (unsigned short)run; run += c; }
        }
        const uint32_t total = s_scan[CVT_BINS - 1];
        __syncthreads();
        for (uint32_t j = 0; j < J; ++j) {                                       // s_rank is free now: it becomes the sorted list of sample ids
            const uint32_t idx = s_choice[j * CVT_BINS + tid];
            if (idx != 0xffffu) {
                const unsigned int old = atomicAdd(reinterpret_cast<unsigned int*>(s_hist) + (idx >> 1), (idx & 1u) ? 0x10000u : 1u);
                const uint32_t p = (idx & 1u) ? (old >> 16) : (old & 0xffffu);
                s_rank[p] = (unsigned short)(j * CVT_BINS + tid);
            }
        }
        __syncthreads();
        // 4. the samples, in list order.  A warp takes 32 consecutive sorted samples; they belong to a handful of consecutive list entries
        //    (~14 samples per region at BASELINE config 4), whose interpolation samples and boxes the warp first copies into its own
        //    shared-memory slots (coalesced 1 KB reads, once per region and warp) — the evaluation then reads shared memory, where lanes on
        //    the same region broadcast and lanes on different regions hit different banks (odd slot stride), instead of issuing 243 global
        //    loads that each touch as many cache lines as the warp has regions.
        const uint64_t slot_pass = slot_tile + uint64_t(j0) * CVT_BINS;
        constexpr int SD = R::ipow(S, D), DW = CvSlot<S, D>::DW, SLOT = cvt_slot_words(DW, D);            // data, rmin, rmax; stride = 4 (mod 8) words: 16-byte aligned records whose
                                                                                  // 128-bit reads fall into different banks for the CVT_SLOTS records of a warp
        float* s_slots = s_region + (tid >> 5) * (CVT_SLOTS * SLOT);
        const uint32_t lane = tid & 31u;
        for (uint32_t pbase = (tid >> 5) * 32u; pbase < J * CVT_BINS; pbase += CVT_BINS) {
            const uint32_t pp = pbase + lane;
            const uint64_t slot = slot_pass + pp;
            const bool valid = pp < total;
            uint32_t sid = 0, idx = 0xffffffffu;
            if (valid) { sid = s_rank[pp]; idx = s_choice[sid]; }
            else {
                a.owner[slot] = 0xffffu; a.weight[slot] = 0.0f; a.app[slot] = 0.0f;
#pragma unroll
                for (int d = 0; d < D; ++d) a.points[uint64_t(d) * a.slots + slot] = a.dom.rmin[d];
            }
            const uint32_t prev = __shfl_up_sync(0xffffffffu, idx, 1);
            const bool is_head = valid && (lane == 0 || idx != prev);
            const unsigned heads = __ballot_sync(0xffffffffu, is_head);
            const int k = __popc(heads);
            const int mine = __popc(heads & (0xffffffffu >> (31 - lane))) - 1;       // which of the warp's regions this lane's sample uses
            const uint32_t rid = valid ? a.list[lo + idx] : 0u;                      // every lane looks its own region up: one round trip, not one per region
            constexpr int NV = (SD + 31) / 32;
            for (int r0 = 0; r0 < k; r0 += CVT_SLOTS) {
                __syncwarp();
                // all of the round's loads first, then the stores: one memory round trip per round of CVT_SLOTS regions
                float v[CVT_SLOTS][NV], rg[CVT_SLOTS];
                unsigned have = 0;
#pragma unroll
                for (int q = 0; q < CVT_SLOTS; ++q) {
                    const unsigned hm = __ballot_sync(0xffffffffu, is_head && mine == r0 + q);      // the lane that heads region r0 + q, if there is one
                    const uint64_t r = __shfl_sync(0xffffffffu, rid, hm ? __ffs(int(hm)) - 1 : 0);
                    if (hm) {
                        have |= 1u << q;
#pragma unroll
                        for (int t = 0; t < NV; ++t) if (t * 32 + int(lane) < SD) v[q][t] = __ldg(a.aos + r * uint64_t(SD) + t * 32 + lane);
                        if (lane < 2 * D) rg[q] = __ldg(a.ranges + r * uint64_t(2 * D) + lane);
                    }
                }
#pragma unroll
                for (int q = 0; q < CVT_SLOTS; ++q) {
                    if (have & (1u << q)) {
                        float* dst = s_slots + q * SLOT;
#pragma unroll
                        for (int t = 0; t < NV; ++t) if (t * 32 + int(lane) < SD) dst[CvSlot<S, D>::pos(t * 32 + int(lane))] = v[q][t];
                        if (lane < 2 * D) dst[DW + lane] = rg[q];
                    }
                }
                __syncwarp();
                if (valid && mine >= r0 && mine < r0 + CVT_SLOTS) {
                    const float* reg = s_slots + (mine - r0) * SLOT;
                    const uint32_t j = sid / CVT_BINS, bt = sid % CVT_BINS;
                    uint32_t bp[2]; bp[0] = o[0] + bt % g.tile[0]; bp[1] = o[1] + bt / g.tile[0];
                    const uint64_t sbin = uint64_t(bp[0]) + uint64_t(bp[1]) * g.res[0];
                    float Lg[D][S]; float vol = 1.0f; u32x4 rnd{0, 0, 0, 0};
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        const float rlo = reg[DW + d], rhi = reg[DW + D + d];
                        float ba = a.dom.rmin[d], bb = a.dom.rmax[d];
                        if (d < 2) { ba = fmaf(float(bp[d]), a.dom.drange[d], a.dom.rmin[d]); bb = fmaf(float(bp[d] + 1u), a.dom.drange[d], a.dom.rmin[d]); }
                        const float ia = fmaxf(ba, rlo), ib = fmaxf(ia, fminf(bb, rhi)), wd = ib - ia;
                        vol *= wd;
                        // a float coordinate takes 24 bits: one Philox block feeds five of them (the top 24 bits of each word + the four low bytes)
                        if (d % 5 == 0) rnd = philox4x32<10>(u32x4{uint32_t(sbin), uint32_t(sbin >> 32), j0 + j, uint32_t(1 + d / 5)}, a.k0, a.k1);
                        const uint32_t u = d % 5 == 0 ? rnd.x : d % 5 == 1 ? rnd.y : d % 5 == 2 ? rnd.z : d % 5 == 3 ? rnd.w
                                         : (((rnd.x & 255u) << 24) | ((rnd.y & 255u) << 16) | ((rnd.z & 255u) << 8));
                        const float xd = fmaf(viltrum::b200::u01(u), wd, ia);
                        a.points[uint64_t(d) * a.slots + slot] = xd;
                        lagrange_basis<S>(rhi > rlo ? __fdividef(xd - rlo, rhi - rlo) : 0.0f, Lg[d]);
                    }
                    a.weight[slot] = vol;
                    a.app[slot] = fast_eval_pairs<S, D>(reg, Lg);
                    a.owner[slot] = (unsigned short)sid;
                }
            }
        }
        __syncthreads();
    }
}

// one CTA per tile: the pass's (f, interpolant, weight) come back into sample order through shared memory, one thread per bin folds them
__global__ void __launch_bounds__(256) cv_tile_accumulate_kernel(TileGeomCv g, uint64_t begin, uint64_t end, uint64_t tile0, uint64_t nbins_total, uint32_t spp, uint32_t J,
                                                                 const uint32_t* __restrict__ count, const float* __restrict__ approx, uint64_t count_base,
                                                                 const float* __restrict__ fval, const float* __restrict__ app, const float* __restrict__ weight,
                                                                 const unsigned short* __restrict__ owner, float* __restrict__ out, int fixed_weight, double fixed_alpha, uint32_t accpass) {
    extern __shared__ unsigned char cvt_smem[];
    float* s_f = reinterpret_cast<float*>(cvt_smem); float* s_a = s_f + accpass * CVT_BINS; float* s_w = s_a + accpass * CVT_BINS;
    const uint64_t t = tile0 + blockIdx.x;
    const uint32_t tid = threadIdx.x;
    uint32_t pos[2]; pos[0] = uint32_t(t % g.tiles[0]) * g.tile[0] + tid % g.tile[0]; pos[1] = uint32_t(t / g.tiles[0]) * g.tile[1] + tid / g.tile[0];
    const uint64_t bin = uint64_t(pos[0]) + uint64_t(pos[1]) * g.res[0];
    const bool live = pos[0] < g.res[0] && pos[1] < g.res[1] && bin >= begin && bin < end;
    const uint32_t cnt = live ? count[bin - count_base] : 0u;
    const double factor = double(nbins_total), rr_uniform = double(cnt);
    CvAccumulator acc;
    const uint64_t slot_tile = uint64_t(blockIdx.x) * CVT_BINS * spp;
    for (uint32_t j0 = 0; j0 < spp; j0 += J) {
        const uint32_t Jp = min(J, spp - j0);
        const uint64_t slot_pass = slot_tile + uint64_t(j0) * CVT_BINS;
        for (uint32_t h0 = 0; h0 < Jp; h0 += accpass) {      // the pass's samples h0 .. h0+accpass-1 of every bin at a time (3 KB of shared memory per sample row)
            const uint32_t Jh = min(accpass, Jp - h0);
            __syncthreads();
            for (uint32_t pb = tid; pb < Jp * CVT_BINS; pb += 8 * CVT_BINS) {       // eight slots per thread at a time: the loads of a batch are in flight together
                uint32_t sid[8]; float vf[8], va[8], vw[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { const uint32_t p = pb + u * CVT_BINS; sid[u] = p < Jp * CVT_BINS ? owner[slot_pass + p] : 0xffffu; }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t p = pb + u * CVT_BINS, j = sid[u] / CVT_BINS;
                    const bool take = sid[u] != 0xffffu && j >= h0 && j < h0 + Jh;
                    sid[u] = take ? sid[u] - h0 * CVT_BINS : 0xffffffffu;
                    if (take) { vf[u] = fval[slot_pass + p]; va[u] = app[slot_pass + p]; vw[u] = weight[slot_pass + p]; }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) if (sid[u] != 0xffffffffu) { s_f[sid[u]] = vf[u]; s_a[sid[u]] = va[u]; s_w[sid[u]] = vw[u]; }
            }
            __syncthreads();
            if (cnt > 0u) for (uint32_t j = 0; j < Jh; ++j) acc.push(s_f[j * CVT_BINS + tid], s_a[j * CVT_BINS + tid], s_w[j * CVT_BINS + tid], factor, rr_uniform, fixed_alpha);
        }
    }
    if (live) out[bin] = acc.result(approx[bin - count_base], fixed_weight, fixed_alpha);
}

__global__ void transpose_chosen_kernel(uint64_t nb, uint32_t spp, const uint32_t* __restrict__ in /* [bin][spp] */, uint32_t* __restrict__ out /* [spp][bin] */) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nb * spp) return;
    const uint64_t b = i % nb; const uint32_t j = uint32_t(i / nb);
    out[i] = in[b * spp + j];
}

template<int S, int D>
int launch_samples(vb200_ctx* ctx, bool replay, bool fast, const vb200_domain& dom, uint64_t begin, uint64_t nb, uint32_t spp, uint32_t k0, uint32_t k1,
                   const vb200_regions* r, const float* aos, const uint32_t* sreg, const uint32_t* sidx, const float* replay_points, float* points, float* weight, float* app) {
    const uint64_t N = nb * spp;
    const unsigned grid = unsigned((N + 127) / 128);
    if (replay) cv_samples_kernel<S, D, true><<<grid, 128, 0, ctx->stream>>>(dom, begin, nb, spp, k0, k1, r->capacity, r->rmin, r->rmax, aos, sreg, sidx, replay_points, points, weight, app);
    else if (fast) cv_samples_kernel<S, D, false, true><<<grid, 128, 0, ctx->stream>>>(dom, begin, nb, spp, k0, k1, r->capacity, r->rmin, r->rmax, aos, sreg, sidx, replay_points, points, weight, app);
    else cv_samples_kernel<S, D, false><<<grid, 128, 0, ctx->stream>>>(dom, begin, nb, spp, k0, k1, r->capacity, r->rmin, r->rmax, aos, sreg, sidx, replay_points, points, weight, app);
    ctx->launches++;
    VB200_CUDA(ctx, cudaGetLastError());
    return VB200_OK;
}

int dispatch_samples(vb200_ctx* ctx, bool replay, bool fast, const vb200_domain& dom, uint64_t begin, uint64_t nb, uint32_t spp, uint32_t k0, uint32_t k1,
                     const vb200_regions* r, const float* aos, const uint32_t* sreg, const uint32_t* sidx, const float* replay_points, float* points, float* weight, float* app) {
#define VB200_CVS(SS, DD) if (r->SH == SS && r->dim == DD) return launch_samples<SS, DD>(ctx, replay, fast, dom, begin, nb, spp, k0, k1, r, aos, sreg, sidx, replay_points, points, weight, app);
    VB200_CVS(3, 1) VB200_CVS(3, 2) VB200_CVS(3, 3) VB200_CVS(3, 4) VB200_CVS(3, 5) VB200_CVS(3, 6)
    VB200_CVS(5, 1) VB200_CVS(5, 2) VB200_CVS(5, 3) VB200_CVS(5, 4)
    VB200_CVS(2, 1) VB200_CVS(2, 2) VB200_CVS(2, 3) VB200_CVS(2, 4) VB200_CVS(2, 5) VB200_CVS(2, 6)
#undef VB200_CVS
    return fail(ctx, VB200_ERR_UNSUPPORTED, "control variates: no kernel for %d samples per dimension in %d dimensions", r->SH, r->dim);
}

template<int D>
int launch_importance(vb200_ctx* ctx, int rs, double power, double cutoff, const vb200_domain& dom, uint64_t begin, uint64_t nb, uint32_t spp, uint32_t k0, uint32_t k1,
                      const vb200_regions* r, const float* aos, const uint32_t* sreg, const uint32_t* sidx, float* points, float* weight, float* app) {
    const uint64_t N = nb * spp;
    const unsigned grid = unsigned((N + 127) / 128);
    if (rs == VB200_RS_IMPORTANCE) cv_samples_importance_kernel<D, 1><<<grid, 128, 0, ctx->stream>>>(dom, begin, nb, spp, k0, k1, r->capacity, r->rmin, r->rmax, aos,
            sreg, sidx, points, weight, app, power, cutoff);
    else if (rs == VB200_RS_MIS) cv_samples_importance_kernel<D, 2><<<grid, 128, 0, ctx->stream>>>(dom, begin, nb, spp, k0, k1, r->capacity, r->rmin, r->rmax, aos, sreg,
            sidx, points, weight, app, power, cutoff);
    else cv_samples_importance_kernel<D, 3><<<grid, 128, 0, ctx->stream>>>(dom, begin, nb, spp, k0, k1, r->capacity, r->rmin, r->rmax, aos, sreg, sidx, points, weight, app, power, cutoff);
    ctx->launches++;
    VB200_CUDA(ctx, cudaGetLastError());
    return VB200_OK;
}
int dispatch_importance(vb200_ctx* ctx, int rs, double power, double cutoff, const vb200_domain& dom, uint64_t begin, uint64_t nb, uint32_t spp, uint32_t k0, uint32_t k1,
                        const vb200_regions* r, const float* aos, const uint32_t* sreg, const uint32_t* sidx, float* points, float* weight, float* app) {
    switch (r->dim) {
        case 1: return launch_importance<1>(ctx, rs, power, cutoff, dom, begin, nb, spp, k0, k1, r, aos, sreg, sidx, points, weight, app);
        case 2: return launch_importance<2>(ctx, rs, power, cutoff, dom, begin, nb, spp, k0, k1, r, aos, sreg, sidx, points, weight, app);
        case 3: return launch_importance<3>(ctx, rs, power, cutoff, dom, begin, nb, spp, k0, k1, r, aos, sreg, sidx, points, weight, app);
        case 4: return launch_importance<4>(ctx, rs, power, cutoff, dom, begin, nb, spp, k0, k1, r, aos, sreg, sidx, points, weight, app);
        case 5: return launch_importance<5>(ctx, rs, power, cutoff, dom, begin, nb, spp, k0, k1, r, aos, sreg, sidx, points, weight, app);
    }
    return fail(ctx, VB200_ERR_UNSUPPORTED, "importance sampling is instantiated for 1..5 dimensions");
}

struct DevBuf {
    void* p = nullptr; vb200_ctx* owner = nullptr;
    ~DevBuf() { if (owner) dfree(owner, p); }
    int alloc(vb200_ctx* ctx, size_t bytes) { owner = ctx; if (dmalloc(ctx, &p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return fail(ctx, VB200_ERR_NOMEM,
            "cudaMalloc of %zu bytes failed", bytes); } return VB200_OK; }
    template<class T> T* as() const { return static_cast<T*>(p); }
};

template<int S, int D> int launch_tile_samples(vb200_ctx* ctx, const CvTileArgs& a, unsigned ntiles) {
    constexpr int SLOT = cvt_slot_words(CvSlot<S, D>::DW, D);
    const size_t smem = size_t(a.J) * CVT_BINS * 2 * 2 + size_t(CVT_MAXLIST) * 2 + std::max(size_t(CVT_MAXLIST) * 4, size_t(8) * CVT_SLOTS * SLOT * 4);
    auto k = cv_tile_samples_kernel<S, D>;
    VB200_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ctx->ktimer) { VB200_CUDA(ctx, cudaEventCreate(&e0)); VB200_CUDA(ctx, cudaEventCreate(&e1)); VB200_CUDA(ctx, cudaEventRecord(e0, ctx->stream)); }
    k<<<ntiles, 256, smem, ctx->stream>>>(a);
    ctx->launches++;
    if (ctx->ktimer) { VB200_CUDA(ctx, cudaEventRecord(e1, ctx->stream)); ctx->ktimer_events.emplace_back(e0, e1); }
    VB200_CUDA(ctx, cudaGetLastError());
    return VB200_OK;
}
// the slab loop of the tile-major residual pass; VB200_ERR_UNSUPPORTED = no kernel for this (S, D): the caller takes the sample-major pipeline
int cv_tile_run(vb200_ctx* ctx, const vb200_integrand* f, const vb200_regions* r, const vb200_cv_params* p, const vb200_domain& dom, const BinWalk& w,
                uint64_t begin, uint64_t end, uint64_t total, uint32_t spp, const float* aos, const uint32_t* count, const float* approx, float* out) {
    const int D = r->dim;
    if (!((r->SH == 3 && D >= 2 && D <= 6) || (r->SH == 2 && D >= 2 && D <= 6) || (r->SH == 5 && D >= 2 && D <= 4))) return VB200_ERR_UNSUPPORTED;
    const uint64_t res0 = dom.res[0], tiles_x = w.tiles[0];
    const uint64_t row0 = begin / res0, row1 = (end - 1) / res0 + 1;                      // bin rows the shard touches
    // slabs of whole tile rows, <= ~32 Mi sample slots each
    uint64_t tile_rows_per_slab = (32ull << 20) / (uint64_t(spp) * CVT_BINS * tiles_x); if (tile_rows_per_slab < 1) tile_rows_per_slab = 1;
    const uint64_t ty0 = row0 / 16, ty1 = (row1 - 1) / 16 + 1;
    if (tile_rows_per_slab > ty1 - ty0) tile_rows_per_slab = ty1 - ty0;
    const uint64_t slots = tile_rows_per_slab * tiles_x * CVT_BINS * spp;
    if (slots > 0x7fffffffull * 4) return VB200_ERR_UNSUPPORTED;
    DevBuf points, weight, app, fval, owner, ranges;
    int rc;
    if ((rc = ranges.alloc(ctx, r->count * uint64_t(2 * D) * 4))) return rc;
    { const uint64_t n = r->count * uint64_t(2 * D);
      ranges_to_aos_kernel<<<unsigned((n + 255) / 256), 256, 0, ctx->stream>>>(r->count, r->capacity, D, r->rmin, r->rmax, ranges.as<float>()); ctx->launches++; VB200_CUDA(ctx, cudaGetLastError()); }
    if ((rc = points.alloc(ctx, slots * D * 4)) || (rc = weight.alloc(ctx, slots * 4)) || (rc = app.alloc(ctx, slots * 4)) || (rc = fval.alloc(ctx, slots * 4))
            || (rc = owner.alloc(ctx, slots * 2))) return rc;
    uint32_t J = spp < uint32_t(CVT_MAXPASS) ? spp : uint32_t(CVT_MAXPASS);
    if (const char* e = std::getenv("VB200_CVT_J")) { const long v = std::atol(e); if (v >= 1 && v <= CVT_MAXPASS && uint32_t(v) < J) J = uint32_t(v); }      // samples per bin and pass (tuning knob)
    uint32_t accpass = CVT_ACCPASS;
    if (const char* e = std::getenv("VB200_CVT_ACCPASS")) { const long v = std::atol(e); if (v >= 1 && v <= 64) accpass = uint32_t(v); }      // tuning knob
    const size_t smem_acc = size_t(accpass) * CVT_BINS * 4 * 3;
    VB200_CUDA(ctx, cudaFuncSetAttribute(cv_tile_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_acc)));
    for (uint64_t ty = ty0; ty < ty1; ty += tile_rows_per_slab) {
        const uint64_t tye = ty + tile_rows_per_slab < ty1 ? ty + tile_rows_per_slab : ty1;
        const unsigned ntiles = unsigned((tye - ty) * tiles_x);
        CvTileArgs a; std::memset(&a, 0, sizeof(a));
        a.g = make_geom_cv(w, dom); a.dom = dom; a.cap = w.cap; a.begin = begin; a.end = end; a.tile0 = ty * tiles_x; a.nbins_total = total;
        a.spp = spp; a.J = J; a.k0 = uint32_t(p->seed); a.k1 = uint32_t(p->seed >> 32);
        a.pstart = w.pstart; a.pend = w.pend; a.offsets = w.tile_offset; a.list = w.tile_list; a.count = count; a.count_base = begin;
        a.ranges = ranges.as<float>(); a.aos = aos;
        a.points = points.as<float>(); a.weight = weight.as<float>(); a.app = app.as<float>(); a.owner = owner.as<unsigned short>();
        a.slots = uint64_t(ntiles) * CVT_BINS * spp;
        rc = VB200_ERR_UNSUPPORTED;
#define VB200_CVT(SS, DD) if (r->SH == SS && D == DD) rc = launch_tile_samples<SS, DD>(ctx, a, ntiles);
        VB200_CVT(3, 2) VB200_CVT(3, 3) VB200_CVT(3, 4) VB200_CVT(3, 5) VB200_CVT(3, 6)
        VB200_CVT(2, 2) VB200_CVT(2, 3) VB200_CVT(2, 4) VB200_CVT(2, 5) VB200_CVT(2, 6)
        VB200_CVT(5, 2) VB200_CVT(5, 3) VB200_CVT(5, 4)
#undef VB200_CVT
        if (rc) return rc;
        vb200_eval_launch ev; std::memset(&ev, 0, sizeof(ev));
        ev.n = a.slots; ev.dim = D; ev.points = points.as<float>(); ev.values = fval.as<float>();
        rc = call_thunk(ctx, f, VB200_K_EVAL_POINTS, &ev); if (rc) return rc;
        cv_tile_accumulate_kernel<<<ntiles, 256, smem_acc, ctx->stream>>>(a.g, begin, end, a.tile0, total, spp, J, count, approx, begin,
                                                                            fval.as<float>(), app.as<float>(), weight.as<float>(), owner.as<unsigned short>(), out,
                                                                            p->weight_strategy == VB200_CV_FIXED_WEIGHT ? 1 : 0, p->alpha, accpass);
        ctx->launches++;
        VB200_CUDA(ctx, cudaGetLastError());
    }
    VB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // the slab buffers die here
    return VB200_OK;
}

// bins of a shard are processed in slabs so that the per-sample buffers stay bounded (<= ~32 Mi samples at a time)
int cv_run(vb200_ctx* ctx, const vb200_integrand* f, const vb200_regions* r, const vb200_cv_params* p, bool replay,
           const uint32_t* replay_chosen, const float* replay_samples, int replay_mem,
           float* bins, int bins_mem, uint32_t* nregions, float* approx_out) {
    VB200_CUDA(ctx, cudaSetDevice(ctx->device));
    if (f->dim <= 0 || f->dim != r->dim) return fail(ctx, VB200_ERR_INVALID, "integrand takes %d dimensions, regions have %d", f->dim, r->dim);
    if (VB200_RULE_IS_STEPS(r->rule)) return fail(ctx, VB200_ERR_UNSUPPORTED, "control variates over a composite (steps) rule table are not supported");
    if (r->f64 || (f->flags & VB200_INTEGRAND_F64)) return fail(ctx, VB200_ERR_UNSUPPORTED, "control variates are computed in fp32: double region tables / integrands are not supported");
    int rc = check_domain(ctx, p->domain, r->dim); if (rc) return rc;
    if (p->spp > 0xffffffffull) return fail(ctx, VB200_ERR_INVALID, "spp invalid");
    if (p->weight_strategy != VB200_CV_OPTIMIZE_WEIGHT && p->weight_strategy != VB200_CV_FIXED_WEIGHT) return fail(ctx, VB200_ERR_INVALID,
            "unknown control-variate weight strategy %d", p->weight_strategy);
    if (p->rr_policy < VB200_RR_UNIFORM || p->rr_policy > VB200_RR_STRATIFIED) return fail(ctx, VB200_ERR_INVALID, "unknown Russian-roulette policy %d", p->rr_policy);
    if (p->rs_policy < VB200_RS_UNIFORM || p->rs_policy > VB200_RS_RUSSIAN_ROULETTE) return fail(ctx, VB200_ERR_INVALID, "unknown region-sampling policy %d", p->rs_policy);
    if (p->rs_policy != VB200_RS_UNIFORM) {
        if (replay) return fail(ctx, VB200_ERR_UNSUPPORTED, "replay feeds recorded sample points: it implies region_sampling_uniform weights");
        if (r->SH != 3) return fail(ctx, VB200_ERR_UNSUPPORTED, "importance sampling of a region needs a Simpson-based table (Simpson::sample, rules.h:184-247)");
        if (r->dim > 5) return fail(ctx, VB200_ERR_UNSUPPORTED, "importance sampling is instantiated for up to 5 dimensions");
    }
    if (p->rr_policy == VB200_RR_STRATIFIED && replay) return fail(ctx, VB200_ERR_UNSUPPORTED, "replay of the stratified allocation is not supported");
    const int policy = p->rr_policy;
    const bool weighted = policy == VB200_RR_INTEGRAL || policy == VB200_RR_ERROR || policy == VB200_RR_PDF;
    const vb200_domain dom = finish_domain(p->domain);
    const uint64_t total = nbins_of(dom);
    uint64_t begin, end; rc = resolve_shard(ctx, p->shard, total, &begin, &end); if (rc) return rc;
    if (begin == end) return VB200_OK;
    const uint32_t spp = uint32_t(p->spp);
    const int D = r->dim;
    const uint64_t nshard = end - begin;
    BinStage st; rc = stage_bins_in(ctx, bins, bins_mem, begin, end, /*upload=*/false, &st); if (rc) return rc;

    DevBuf aos; rc = aos.alloc(ctx, r->count * uint64_t(r->sd) * sizeof(float)); if (rc) return rc;
    { const uint64_t n = r->count * uint64_t(r->sd);
      regions_to_aos_kernel<<<unsigned((n + 255) / 256), 256, 0, ctx->stream>>>(r->count, r->capacity, r->sd, r->data, aos.as<float>()); ctx->launches++; VB200_CUDA(ctx, cudaGetLastError()); }

    DevBuf d_count, d_approx; rc = d_count.alloc(ctx, nshard * sizeof(uint32_t)); if (rc) return rc; rc = d_approx.alloc(ctx, nshard * sizeof(float)); if (rc) return rc;
    BinWalk w;
    rc = walk_build(ctx, r, dom, begin, end, &w); if (rc) return rc;
    struct WalkGuard { BinWalk* w; ~WalkGuard() { walk_free(w); } } guard{&w};
    // FAST integrands (the throughput path: statistical parity) take the fp32 forms of the two arithmetic-heavy kernels — the bin walk and
    // the interpolant under the residual samples; EXACT integrands and replay keep every rounding of the reference (VB200_CV_EXACT=1 forces that)
    const char* force_exact = std::getenv("VB200_CV_EXACT");
    const bool fast = !replay && !(f->flags & VB200_INTEGRAND_EXACT) && !(force_exact && force_exact[0] == '1');
    if (!(fast && walk_accumulate_fast(ctx, r, w, dom, begin, end, 1, nullptr, d_approx.as<float>(), d_count.as<uint32_t>(), &rc)))
        rc = walk_accumulate(ctx, r, w, dom, begin, end, 1, nullptr, d_approx.as<float>(), d_count.as<uint32_t>());
    if (rc) return rc;
    // weighted roulettes: per-bin sum of the pair weights, then of the clamped weights (two more walks, nothing per pair is stored)
    DevBuf d_wsum, d_csum, d_rerr, d_pdf;
    if (weighted && spp > 0) {
        if (policy == VB200_RR_PDF) { float* pp = nullptr; rc = walk_pdf_patches(ctx, r, w, &pp); d_pdf.p = pp; d_pdf.owner = ctx; if (rc) return rc; }
        if ((rc = d_wsum.alloc(ctx, nshard * sizeof(double))) || (rc = d_csum.alloc(ctx, nshard * sizeof(double)))) return rc;
        if (policy == VB200_RR_ERROR) { if ((rc = d_rerr.alloc(ctx, r->count * sizeof(float))) || (rc = region_total_errors(ctx, r, w, d_rerr.as<float>()))) return rc; }
        for (int pass = 1; pass <= 2; ++pass) {
            rc = walk_rr_pass(ctx, r, w, dom, begin, end, begin, policy, pass, d_rerr.as<float>(), d_pdf.as<float>(), d_count.as<uint32_t>(), d_wsum.as<double>(),
                    d_csum.as<double>(), 0, nullptr, nullptr);
            if (rc) return rc;
        }
    }

    // tile-major residual pass (cv_tile_samples_kernel): the throughput path of the crespo2021 preset over a 2-D bin grid
    const char* tile_env = std::getenv("VB200_CV_TILE");       // test knob: 0 = keep the sample-major pipeline
    const bool tile_path = fast && policy == VB200_RR_UNIFORM && p->rs_policy == VB200_RS_UNIFORM && spp > 0 && w.db == 2 && w.tile[0] == 16 && w.tile[1] == 16
            && w.max_list <= uint64_t(CVT_MAXLIST) &&
                           (r->SH == 2 || r->SH == 3 || r->SH == 5) && !(tile_env && tile_env[0] == '0');
    if (tile_path) {
        rc = cv_tile_run(ctx, f, r, p, dom, w, begin, end, total, spp, aos.as<float>(), d_count.as<uint32_t>(), d_approx.as<float>(), st.dev_base);
        if (rc != VB200_ERR_UNSUPPORTED) { if (rc) return rc; goto residual_done; }
    }
    if (spp > 0) {
        uint64_t slab = (32ull << 20) / spp; if (slab < 1) slab = 1; if (slab > nshard) slab = nshard;
        const uint64_t NS = slab * spp;
        DevBuf rank, chosen, points, weight, app, fval, rchosen, rpoints, rrf, sreg, sidx, iota, rec, sort_tmp;
        if (policy != VB200_RR_UNIFORM) { if ((rc = rrf.alloc(ctx, NS * sizeof(double)))) return rc; }
        // region-sorted visiting order of the residual samples (cv_samples_kernel): radix sort of (region, sample position) pairs
        if (NS > 0x7fffffffull) return fail(ctx, VB200_ERR_UNSUPPORTED, "control variates: slab of %llu samples too large", (unsigned long long)NS);
        if ((rc = sreg.alloc(ctx, NS * 4)) || (rc = sidx.alloc(ctx, NS * 4)) || (rc = iota.alloc(ctx, NS * 4)) || (rc = rec.alloc(ctx, NS * sizeof(float4)))) return rc;
        int region_bits = 1; while ((1ull << region_bits) < r->count && region_bits < 32) ++region_bits;
        size_t sort_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, chosen.as<uint32_t>(), sreg.as<uint32_t>(), iota.as<uint32_t>(), sidx.as<uint32_t>(), int(NS), 0, region_bits, ctx->stream);
        if ((rc = sort_tmp.alloc(ctx, sort_bytes))) return rc;
        cv_iota_kernel<<<unsigned((NS + 255) / 256), 256, 0, ctx->stream>>>(NS, iota.as<uint32_t>()); ctx->launches++; VB200_CUDA(ctx, cudaGetLastError());
        if ((rc = chosen.alloc(ctx, NS * 4)) || (rc = points.alloc(ctx, NS * D * 4)) || (rc = weight.alloc(ctx, NS * 4)) || (rc = app.alloc(ctx, NS * 4)) || (rc = fval.alloc(ctx, NS * 4))) return rc;
        if (!replay) { if ((rc = rank.alloc(ctx, NS * 4))) return rc; }
        else if (replay_mem == VB200_HOST) { if ((rc = rchosen.alloc(ctx, NS * 4)) || (rc = rpoints.alloc(ctx, NS * D * 4))) return rc; }
        else { if ((rc = rchosen.alloc(ctx, 4))) return rc; }
        const TileGeomCv g = make_geom_cv(w, dom);
        for (uint64_t s0 = begin; s0 < end; s0 += slab) {
            const uint64_t s1 = s0 + slab < end ? s0 + slab : end, nb = s1 - s0, N = nb * spp;
            const uint32_t* cnt = d_count.as<uint32_t>() + (s0 - begin);
            const float* rp = nullptr;
            if (!replay && weighted) {
                cv_raw_kernel<<<unsigned((N + 255) / 256), 256, 0, ctx->stream>>>(s0, nb, spp, uint32_t(p->seed), uint32_t(p->seed >> 32), rank.as<uint32_t>());
                ctx->launches++; VB200_CUDA(ctx, cudaGetLastError());
                VB200_CUDA(ctx, cudaMemsetAsync(chosen.p, 0, N * 4, ctx->stream));
                rc = walk_rr_pass(ctx, r, w, dom, s0, s1, begin, policy, 3, d_rerr.as<float>(), d_pdf.as<float>(), d_count.as<uint32_t>(), d_wsum.as<double>(), d_csum.as<double>(), spp,
                                  rank.as<uint32_t>(), chosen.as<uint32_t>()); if (rc) return rc;
            } else if (!replay) {
                if (policy == VB200_RR_STRATIFIED) cv_stratified_ranks_kernel<<<unsigned((N + 255) / 256), 256, 0, ctx->stream>>>(s0, nb, spp, uint32_t(p->seed),
                        uint32_t(p->seed >> 32), cnt, rank.as<uint32_t>(), rrf.as<double>());
                else cv_ranks_kernel<<<unsigned((N + 255) / 256), 256, 0, ctx->stream>>>(s0, nb, spp, uint32_t(p->seed), uint32_t(p->seed >> 32), cnt, rank.as<uint32_t>());
                ctx->launches++; VB200_CUDA(ctx, cudaGetLastError());
                VB200_CUDA(ctx, cudaMemsetAsync(chosen.p, 0, N * 4, ctx->stream));
                const char* legacy = std::getenv("VB200_CV_RESOLVE_LEGACY");       // test knob: the chunk-by-chunk kernel
                if (legacy && legacy[0] == '1') {
                    if (w.db == 1) cv_resolve_kernel<1><<<unsigned(w.ntiles), 256, 0, ctx->stream>>>(g, w.cap, s0, s1, spp, w.pstart, w.pend, w.tile_offset, w.tile_list,
                            rank.as<uint32_t>(), chosen.as<uint32_t>());
                    else if (w.db == 2) cv_resolve_kernel<2><<<unsigned(w.ntiles), 256, 0, ctx->stream>>>(g, w.cap, s0, s1, spp, w.pstart, w.pend, w.tile_offset,
                            w.tile_list, rank.as<uint32_t>(), chosen.as<uint32_t>());
                    else cv_resolve_kernel<3><<<unsigned(w.ntiles), 256, 0, ctx->stream>>>(g, w.cap, s0, s1, spp, w.pstart, w.pend, w.tile_offset, w.tile_list,
                            rank.as<uint32_t>(), chosen.as<uint32_t>());
                } else {
                    if (w.db == 1) cv_resolve_grouped_kernel<1, 16><<<unsigned(w.ntiles), 256, 0, ctx->stream>>>(g, w.cap, s0, s1, spp, w.pstart, w.pend, w.tile_offset,
                            w.tile_list, rank.as<uint32_t>(), chosen.as<uint32_t>());
                    else if (w.db == 2) cv_resolve_grouped_kernel<2, 16><<<unsigned(w.ntiles), 256, 0, ctx->stream>>>(g, w.cap, s0, s1, spp, w.pstart, w.pend,
                            w.tile_offset, w.tile_list, rank.as<uint32_t>(), chosen.as<uint32_t>());
                    else cv_resolve_grouped_kernel<3, 16><<<unsigned(w.ntiles), 256, 0, ctx->stream>>>(g, w.cap, s0, s1, spp, w.pstart, w.pend, w.tile_offset,
                            w.tile_list, rank.as<uint32_t>(), chosen.as<uint32_t>());
                }
                ctx->launches++; VB200_CUDA(ctx, cudaGetLastError());
            } else {
                const uint32_t* src_c = replay_chosen + (s0 - begin) * spp; const float* src_p = replay_samples + (s0 - begin) * spp * D;
                if (replay_mem == VB200_HOST) {
                    VB200_CUDA(ctx, cudaMemcpyAsync(rchosen.p, src_c, N * 4, cudaMemcpyHostToDevice, ctx->stream));
                    VB200_CUDA(ctx, cudaMemcpyAsync(rpoints.p, src_p, N * D * 4, cudaMemcpyHostToDevice, ctx->stream));
                    src_c = rchosen.as<uint32_t>(); src_p = rpoints.as<float>();
                }
                transpose_chosen_kernel<<<unsigned((N + 255) / 256), 256, 0, ctx->stream>>>(nb, spp, src_c, chosen.as<uint32_t>());
                ctx->launches++; VB200_CUDA(ctx, cudaGetLastError());
                rp = src_p;
            }
            { size_t tb = sort_bytes;
              VB200_CUDA(ctx, cub::DeviceRadixSort::SortPairs(sort_tmp.p, tb, chosen.as<uint32_t>(), sreg.as<uint32_t>(), iota.as<uint32_t>(), sidx.as<uint32_t>(),
                      int(N), 0, region_bits, ctx->stream));
              ctx->launches += 1 + (region_bits + 7) / 8; }
            if (p->rs_policy != VB200_RS_UNIFORM)
                rc = dispatch_importance(ctx, p->rs_policy, p->rs_power, p->rs_cutoff, dom, s0, nb, spp, uint32_t(p->seed), uint32_t(p->seed >> 32), r, aos.as<float>(),
                        sreg.as<uint32_t>(), sidx.as<uint32_t>(),
                                         points.as<float>(), weight.as<float>(), app.as<float>());
            else
                rc = dispatch_samples(ctx, replay, fast, dom, s0, nb, spp, uint32_t(p->seed), uint32_t(p->seed >> 32), r, aos.as<float>(), sreg.as<uint32_t>(), sidx.as<uint32_t>(), rp,
                                      points.as<float>(), weight.as<float>(), app.as<float>());
            if (rc) return rc;
            if (weighted) {
                rc = rr_factors(ctx, r, w, dom, s0, nb, begin, policy, spp, d_rerr.as<float>(), d_pdf.as<float>(), d_count.as<uint32_t>(), d_wsum.as<double>(), d_csum.as<double>(),
                                chosen.as<uint32_t>(), rrf.as<double>()); if (rc) return rc;
            }
            vb200_eval_launch ev; std::memset(&ev, 0, sizeof(ev));
            ev.n = N; ev.dim = D; ev.points = points.as<float>(); ev.values = fval.as<float>();
            rc = call_thunk(ctx, f, VB200_K_EVAL_POINTS, &ev); if (rc) return rc;
            cv_unsort_kernel<<<unsigned((N + 255) / 256), 256, 0, ctx->stream>>>(N, sidx.as<uint32_t>(), fval.as<float>(), app.as<float>(), weight.as<float>(), rec.as<float4>());
            ctx->launches++; VB200_CUDA(ctx, cudaGetLastError());
            cv_accumulate_kernel<<<unsigned((nb + 127) / 128), 128, 0, ctx->stream>>>(s0, nb, spp, total, cnt, d_approx.as<float>() + (s0 - begin),
                                                                                        rec.as<float4>(), st.dev_base,
                                                                                        p->weight_strategy == VB200_CV_FIXED_WEIGHT ? 1 : 0, p->alpha,
                                                                                        policy != VB200_RR_UNIFORM ? rrf.as<double>() : nullptr);
            ctx->launches++; VB200_CUDA(ctx, cudaGetLastError());
        }
        VB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // the slab buffers die here
    } else {
        // no residual samples: bins = approximation (weight-strategy.h:95, size < 2)
        VB200_CUDA(ctx, cudaMemcpyAsync(st.dev_base + begin, d_approx.p, nshard * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    }
residual_done:
    auto copy_out = [&] (void* dst, const void* src, size_t bytes) -> int {
        VB200_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, bins_mem == VB200_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, ctx->stream));
        return VB200_OK;
    };
    if (nregions && (rc = copy_out(nregions, d_count.p, nshard * sizeof(uint32_t)))) return rc;
    if (approx_out && (rc = copy_out(approx_out, d_approx.p, nshard * sizeof(float)))) return rc;
    rc = stage_bins_out(ctx, st); if (rc) return rc;
    VB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VB200_OK;
}

} // namespace

extern "C" int vb200_cv_integrate(vb200_ctx* ctx, const vb200_integrand* f, const vb200_regions* r, const vb200_cv_params* p,
                                  float* bins, int bins_mem, uint32_t* nregions, float* approx) {
    if (!ctx || !f || !r || !p || !bins) return fail(ctx, VB200_ERR_INVALID, "NULL argument");
    return cv_run(ctx, f, r, p, false, nullptr, nullptr, VB200_HOST, bins, bins_mem, nregions, approx);
}

extern "C" int vb200_cv_replay(vb200_ctx* ctx, const vb200_integrand* f, const vb200_regions* r, const vb200_cv_params* p,
                               const uint32_t* chosen, const float* samples, int mem, float* bins, int bins_mem) {
    if (!ctx || !f || !r || !p || !bins || !chosen || !samples) return fail(ctx, VB200_ERR_INVALID, "NULL argument");
    return cv_run(ctx, f, r, p, true, chosen, samples, mem, bins, bins_mem, nullptr, nullptr);
}
