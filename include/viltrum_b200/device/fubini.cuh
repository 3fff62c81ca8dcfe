// Fubini adapters — the device form of the reference's
//   function_split_and_integrate_at<N>(f, monte_carlo(m, seed), range_rest)        src/combination/fubini.h:51-75
// i.e. the N-dimensional integrand  g(x) = vol(rest)/m * sum_{s<m} f(x (+) r_s),  r_s uniform in the rest of the range — finite
// (Range<Float,DIM-N>, f over std::array<float,DIM>) or infinite (RangeInfinite, f over a lazy sequence, concat.h:9-45).
// The adapters are ordinary trivially-copyable functors, so everything that takes an integrand takes them unchanged:
//   integrator_fubini<N>(first, monte_carlo(m))                  -> first integrator over g                (fubini.h:78-101)
//   regions_generator_fubini<N>(adaptive heap, monte_carlo(m))   -> region table of g                      (regions-generator-fubini.h:7-28)
//   residual pass of the control variates                        -> g with m = 1                           (regions-integrator-parallel-variance-reduction.h:69)
// The reference threads ONE mt19937 through all evaluations of g; a GPU wants a stateless stream per evaluation: the stream
// of an evaluation is keyed by the evaluation point itself (Philox of the coordinates' bits), sample s / draw block b of that
// evaluation is Philox(key; stream, s, b).  Evaluating g twice at the same point therefore returns the same estimate — the
// generators evaluate every grid point once and residual samples are continuous random points, so nothing relies on it.
#pragma once
#include <array>
#include <cstring>
#include <cuda_runtime.h>
#include "../../viltrum_b200.h"
#include "philox.cuh"

namespace viltrum { namespace b200 {

VB200_HD uint32_t float_bits(float v) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(v);
#else
    uint32_t u; std::memcpy(&u, &v, sizeof(u)); return u;
#endif
}

// 64-bit stream id of an evaluation point: Philox as a hash, three coordinates absorbed per call
template<int N>
VB200_HD void fubini_stream(const std::array<float,N>& x, uint32_t k0, uint32_t k1, uint32_t& s0, uint32_t& s1) {
    u32x4 c{0x46756269u, 0x6e692d31u, 0u, 0u};
#pragma unroll
    for (int i = 0; i < N; i += 3) {
        c = philox4x32<10>(u32x4{c.x ^ float_bits(x[i]), c.y ^ (i + 1 < N ? float_bits(x[i + 1 < N ? i + 1 : i]) : 0u),
                                 c.z ^ (i + 2 < N ? float_bits(x[i + 2 < N ? i + 2 : i]) : 0u), c.w + uint32_t(i)}, k0, k1);
    }
    s0 = c.x; s1 = c.y;
}

// finite rest: f over std::array<float,N+R>
template<class F, int N, int R>
struct FubiniFinite {
    F f;
    float rmin[R], rext[R];          // rest range: min and (max-min)
    double factor;                   // vol(rest)/m      (monte-carlo.h:43-45 with one bin)
    uint32_t m, k0, k1;
    __host__ __device__ float operator()(const std::array<float,N>& x) const {
        uint32_t s0, s1; fubini_stream<N>(x, k0, k1, s0, s1);
        std::array<float,N+R> full;
#pragma unroll
        for (int i = 0; i < N; ++i) full[i] = x[i];
        float sol = 0.0f;
        for (uint32_t s = 0; s < m; ++s) {
#pragma unroll
            for (int blk = 0; blk < (R + 3) / 4; ++blk) {
                const u32x4 r = philox4x32<10>(u32x4{s0, s1, s, uint32_t(blk)}, k0 ^ 0x9E3779B9u, k1 ^ 0x85EBCA6Bu);
                const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) { const int i = blk * 4 + j; if (i < R) full[N + i] = u01(w[j]) * rext[i] + rmin[i]; }
            }
            sol = float(double(sol) + double(f(full)) * factor);        // T += T*double (monte-carlo.h:59)
        }
        return sol;
    }
};

// infinite rest: f over a sequence; the sequence handed to f is concat(x, random sequence over the rest range)
template<class F, int N>
struct FubiniInfinite {
    F f;
    int nrest;                                           // explicit entries of the rest range (implicit [0,1] beyond, range-infinite.h:31-37)
    float rmin[VB200_MAX_DIM], rext[VB200_MAX_DIM];
    double factor;                                       // vol(rest)/m      (monte-carlo.h:70-72)
    uint32_t m, k0, k1;

    struct Sequence {
        const FubiniInfinite* a; const float* x; uint32_t s0, s1, s;
        class const_iterator {
            const Sequence* q; uint32_t i; u32x4 blk; float n;
            __host__ __device__ void load() {
                if (i < uint32_t(N)) { n = q->x[i]; return; }
                const uint32_t j = i - uint32_t(N);
                if ((j & 3u) == 0u) blk = philox4x32<10>(u32x4{q->s0, q->s1, q->s, j >> 2}, q->a->k0 ^ 0x9E3779B9u, q->a->k1 ^ 0x85EBCA6Bu);
                const float u = pick(blk, int(j & 3u));
                n = int(j) < q->a->nrest ? u * q->a->rext[j] + q->a->rmin[j] : u;
            }
        public:
            __host__ __device__ const_iterator(const Sequence* q_) : q(q_), i(0), blk{0, 0, 0, 0}, n(0.0f) { load(); }
            __host__ __device__ const float& operator*() const { return n; }
            __host__ __device__ const_iterator& operator++() { ++i; load(); return *this; }
            __host__ __device__ bool operator!=(const const_iterator&) const { return true; }
            __host__ __device__ bool operator==(const const_iterator&) const { return false; }
        };
        __host__ __device__ const_iterator begin() const { return const_iterator(this); }
        __host__ __device__ const_iterator end() const { return const_iterator(this); }
    };

    __host__ __device__ float operator()(const std::array<float,N>& x) const {
        Sequence seq; seq.a = this; seq.x = x.data();
        fubini_stream<N>(x, k0, k1, seq.s0, seq.s1);
        float sol = 0.0f;
        for (uint32_t s = 0; s < m; ++s) {
            seq.s = s;
            sol = float(double(sol) + double(f(seq)) * factor);          // monte-carlo.h:81
        }
        return sol;
    }
};

// factories: rest_min/rest_max have `nrest` entries (finite: nrest == R)
template<int N, int R, class F>
inline FubiniFinite<F,N,R> make_fubini_finite(const F& f, const float* rest_min, const float* rest_max, uint64_t mc_samples, uint64_t seed) {
    FubiniFinite<F,N,R> g{f, {}, {}, 0.0, uint32_t(mc_samples), uint32_t(seed), uint32_t(seed >> 32)};
    float vol = 1.0f;
    for (int i = 0; i < R; ++i) { g.rmin[i] = rest_min[i]; g.rext[i] = rest_max[i] - rest_min[i]; vol *= (rest_max[i] - rest_min[i]); }
    g.factor = double(vol) / double(mc_samples);
    return g;
}
template<int N, class F>
inline FubiniInfinite<F,N> make_fubini_infinite(const F& f, const float* rest_min, const float* rest_max, int nrest, uint64_t mc_samples, uint64_t seed) {
    FubiniInfinite<F,N> g{f, nrest < VB200_MAX_DIM ? nrest : VB200_MAX_DIM, {}, {}, 0.0, uint32_t(mc_samples), uint32_t(seed), uint32_t(seed >> 32)};
    float vol = 1.0f;
    for (int i = 0; i < VB200_MAX_DIM; ++i) { g.rmin[i] = 0.0f; g.rext[i] = 1.0f; }
    for (int i = 0; i < g.nrest; ++i) { g.rmin[i] = rest_min[i]; g.rext[i] = rest_max[i] - rest_min[i]; vol *= (rest_max[i] - rest_min[i]); }
    g.factor = double(vol) / double(mc_samples);
    return g;
}

}} // namespace viltrum::b200
