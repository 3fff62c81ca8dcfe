// TEST INFRASTRUCTURE — not product code.  Builds into oracle/liboracle.so (vo_kind() == "port").
//
// Plain CPU restatement of the reference's per-bin integration hot path, written as
// straight loops over flat arrays with run-time dimensions: no templates over rules/regions, no lazy
// expression views.  Every function cites the reference file:line it restates.  It exists to be the
// checker the CUDA path is compared against on machines where /root/reference is absent (the GPU box);
// it is itself pinned bit-for-bit against oracle/_ref/libviltrum_ref.so (the unmodified reference) by
// tests/test_oracle_vs_reference.py and against tests/golden/*.json (vectors generated from the reference).
//
// Third-party arithmetic the reference takes from the C++ standard library (libstdc++ 13, GCC 13.3 — the
// reference pins no version) is restated here explicitly so the oracle does not depend on which standard
// library it is built with:
//   * std::mt19937                     — [rand.eng.mers] MT19937, Matsumoto & Nishimura 1998
//   * std::uniform_real_distribution   — libstdc++ bits/random.h operator(): canonical*(b-a)+a,
//     std::generate_canonical<float,24>  bits/random.tcc:3349-3381: float(u32)/2^32, clamped to nextafter(1,0)
//   * std::uniform_int_distribution    — libstdc++ bits/uniform_int_dist.h:240-330 (Lemire nearly-divisionless)
//   * std::push_heap / std::pop_heap   — libstdc++ bits/stl_heap.h:135-267
//
// Numerics: built with -ffp-contract=off; float/double promotions mirror the reference expression by
// expression — that is what makes the bit-exact parity modes possible.
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include <array>
#include <thread>
#include <algorithm>
#include <functional>
#include <memory>
#include <cstdlib>
#include <cstdio>
#include "integrands.h"
#include "oracle_api.h"

namespace {

// ---------------------------------------------------------------------------------------------------------
// RNG restatement
// ---------------------------------------------------------------------------------------------------------
struct MT19937 {
    uint32_t x[624]; int p;
    explicit MT19937(uint64_t sd = 5489u) { seed(sd); }
    void seed(uint64_t sd) {                       // seed taken mod 2^32 (bits/random.tcc mersenne_twister_engine::seed)
        x[0] = uint32_t(sd);
        for (int i=1;i<624;++i) x[i] = 1812433253u*(x[i-1]^(x[i-1]>>30)) + uint32_t(i);
        p = 624;
    }
    uint32_t operator()() {
        if (p >= 624) {
            for (int k=0;k<624;++k) {
                uint32_t y = (x[k]&0x80000000u) | (x[(k+1)%624]&0x7fffffffu);
                x[k] = x[(k+397)%624] ^ (y>>1) ^ ((y&1u)?0x9908b0dfu:0u);
            }
            p = 0;
        }
        uint32_t z = x[p++];
        z ^= (z>>11); z ^= (z<<7)&0x9d2c5680u; z ^= (z<<15)&0xefc60000u; z ^= (z>>18);
        return z;
    }
};

inline float canonical(MT19937& g) {              // generate_canonical<float,24>, one 32-bit draw
    float r = float(g()) / 4294967296.0f;
    if (r >= 1.0f) r = std::nextafter(1.0f, 0.0f);
    return r;
}
inline float uniform_real(MT19937& g, float a, float b) { return canonical(g)*(b-a)+a; }
inline double canonical_double(MT19937& g) {      // generate_canonical<double,53>, two 32-bit draws (bits/random.tcc:3349-3381)
    double sum = 0.0, tmp = 1.0;
    for (int k=0;k<2;++k) { sum += double(g())*tmp; tmp *= 4294967296.0; }
    double r = sum/tmp;
    if (r >= 1.0) r = std::nextafter(1.0, 0.0);
    return r;
}

inline uint64_t uniform_int(MT19937& g, uint64_t a, uint64_t b) {  // uniform_int_distribution<size_t>(a,b), 32-bit urng
    uint64_t urange = b-a;
    // (the reference only ever asks for ranges far below 2^32: number of regions in a bin)
    uint32_t range = uint32_t(urange+1);
    uint64_t product = uint64_t(g())*uint64_t(range);
    uint32_t low = uint32_t(product);
    if (low < range) {
        uint32_t threshold = uint32_t(-range) % range;
        while (low < threshold) { product = uint64_t(g())*uint64_t(range); low = uint32_t(product); }
    }
    return (product>>32) + a;
}

// ---------------------------------------------------------------------------------------------------------
// integrands by name (finite ones evaluate a float[dim] point)
// ---------------------------------------------------------------------------------------------------------
typedef float (*FiniteFn)(const float*);
template<typename F> float call_finite(const float* x) {
    std::array<float,F::dim> a; for (int i=0;i<F::dim;++i) a[i]=x[i];
    return F()(a);
}
struct FiniteIntegrand { const char* name; int dim; FiniteFn fn; };
const FiniteIntegrand FINITE[] = {
    {"x2y2",2,call_finite<vo::X2Y2>}, {"ind2",2,call_finite<vo::Ind2>}, {"cubic1",1,call_finite<vo::Cubic1>},
    {"poly3",3,call_finite<vo::Poly3>}, {"shade4_64",4,call_finite<vo::Shade4<64>>}, {"shade4_16",4,call_finite<vo::Shade4<16>>},
    {"shade5_64",5,call_finite<vo::Shade5<64>>}, {"shade5_16",5,call_finite<vo::Shade5<16>>}, {"smooth_edge2",2,call_finite<vo::SmoothEdge2>},
};
const FiniteIntegrand* find_finite(const char* n) { for (auto& f : FINITE) if (!std::strcmp(f.name,n)) return &f; return nullptr; }

// lazy sequence protocol of reference random-sequence-ref-dis.h:20-38, over a generator callback
struct LazySeq {
    std::function<float()> next;
    struct It { const LazySeq* s; float n; const float& operator*() const { return n; } It& operator++() { n = s->next(); return *this; } };
    It begin() const { return It{this, next()}; }
};
typedef float (*InfFn)(const LazySeq&);
template<typename F> float call_inf(const LazySeq& s) { return F()(s); }
struct InfIntegrand { const char* name; InfFn fn; };
const InfIntegrand INFINITE[] = { {"walk",call_inf<vo::Walk>}, {"decay",call_inf<vo::Decay>} };
const InfIntegrand* find_inf(const char* n) { for (auto& f : INFINITE) if (!std::strcmp(f.name,n)) return &f; return nullptr; }

// ---------------------------------------------------------------------------------------------------------
// bins: tensor layout, multidimensional_range iteration (dim 0 fastest)   reference tensor.h:17-23,
// multidimensional-range.h:33-40
// ---------------------------------------------------------------------------------------------------------
inline uint64_t nbins_of(int db, const uint64_t* res) { uint64_t n=1; for (int i=0;i<db;++i) n*=res[i]; return n; }
inline void unflatten(uint64_t lin, int db, const uint64_t* res, uint64_t* pos) { for (int i=0;i<db;++i) { pos[i]=lin%res[i]; lin/=res[i]; } }

// Range<float,D>::volume  (range.h:21-25): product in float, starting from 1
inline float volume_of(int d, const float* a, const float* b) { float v=1.0f; for (int i=0;i<d;++i) v*=(b[i]-a[i]); return v; }

// bin sub-box of the first `db` dims (monte-carlo-per-bin-parallel.h:45-47,59-61; same expression in every
// per-bin integrator): drange = (max-min)/float(res); [min+float(pos)*drange, min+float(pos+1)*drange]
inline void bin_box(int d, int db, const float* rmin, const float* rmax, const uint64_t* res, const uint64_t* pos, float* a, float* b) {
    for (int i=0;i<d;++i) { a[i]=rmin[i]; b[i]=rmax[i]; }
    for (int i=0;i<db;++i) {
        float drange = (rmax[i]-rmin[i])/float(res[i]);
        a[i] = rmin[i] + float(pos[i])*drange;
        b[i] = rmin[i] + float(pos[i]+1)*drange;
    }
}

} // namespace

extern "C" const char* vo_kind(void) { return "port"; }
extern "C" int vo_integrand_dim(const char* name) {
    if (auto f = find_finite(name)) return f->dim;
    if (find_inf(name)) return -1;
    return 0;
}

// =========================================================================================================
// Monte Carlo
// =========================================================================================================

// monte-carlo-per-bin-parallel.h:41-71
extern "C" int vo_mc_per_bin_parallel(const char* integrand, int dimbins, const uint64_t* res,
                           const float* rmin, const float* rmax, uint64_t spp, uint64_t seed,
                           float* bins, float* rec_samples, double* rec_sum, double* rec_sum2) {
    auto F = find_finite(integrand); if (!F) return -1;
    int D = F->dim; if (dimbins<1 || dimbins>D) return -2;
    uint64_t nb = nbins_of(dimbins,res);
    double factor = volume_of(D,rmin,rmax)/double(spp);                       // :45
    MT19937 master(seed);
    std::vector<uint32_t> perbin_seed(nb);
    for (uint64_t k=0;k<nb;++k) perbin_seed[k] = uint32_t(master());          // :50-54 (sequential, dim-0-fastest)
    for (uint64_t k=0;k<nb;++k) {                                             // :56-70 (bins are independent)
        uint64_t pos[8]; unflatten(k,dimbins,res,pos);
        MT19937 local(perbin_seed[k]);
        float a[8], b[8]; bin_box(D,dimbins,rmin,rmax,res,pos,a,b);
        double s1=0, s2=0;
        for (uint64_t s=0;s<spp;++s) {
            float x[8];
            for (int i=0;i<D;++i) x[i] = uniform_real(local,a[i],b[i]);       // :64-67
            float v = F->fn(x);
            bins[k] = float(double(bins[k]) + double(v)*factor);              // :68  float += float*double
            if (rec_samples) for (int i=0;i<D;++i) rec_samples[(k*spp+s)*D+i]=x[i];
            s1 += double(v); s2 += double(v)*double(v);
        }
        if (rec_sum) rec_sum[k]=s1;
        if (rec_sum2) rec_sum2[k]=s2;
    }
    return 0;
}

// integrator-per-bin-parallel.h:16-35 wrapping monte-carlo.h:39-63 through integrate.h:105-113
extern "C" int vo_per_bin_parallel_mc(const char* integrand, int dimbins, const uint64_t* res,
                           const float* rmin, const float* rmax, uint64_t spp, uint64_t seed,
                           float* bins, float* rec_samples, double* rec_sum, double* rec_sum2) {
    auto F = find_finite(integrand); if (!F) return -1;
    int D = F->dim; if (dimbins<1 || dimbins>D) return -2;
    uint64_t nb = nbins_of(dimbins,res);
    // monte_carlo(spp,seed) holds mt19937(seed); wrapping it copies it once (copy ctor reseeds from the source
    // stream, monte-carlo.h:32-33; integrator-per-bin-parallel.h:37-38), then one more reseeding copy per bin
    // into tensor<Integrator> in tensor order (integrator-per-bin-parallel.h:24).
    MT19937 user(seed);
    MT19937 master{uint64_t(user())};
    std::vector<uint32_t> perbin_seed(nb);
    for (uint64_t k=0;k<nb;++k) perbin_seed[k] = master();
    for (uint64_t k=0;k<nb;++k) {
        uint64_t pos[8]; unflatten(k,dimbins,res,pos);
        MT19937 local(perbin_seed[k]);
        float a[8], b[8]; bin_box(D,dimbins,rmin,rmax,res,pos,a,b);
        double factor = 1.0*volume_of(D,a,b)/double(spp);                     // monte-carlo.h:43-45 with one bin
        float sol = 0.0f;                                                     // integrate.h:108
        double s1=0, s2=0;
        for (uint64_t s=0;s<spp;++s) {
            float x[8];
            for (int i=0;i<D;++i) x[i] = uniform_real(local,a[i],b[i]);       // monte-carlo.h:50-53
            // is_inside (monte-carlo.h:54) is always true for points drawn inside [a,b]
            float v = F->fn(x);
            sol = float(double(sol) + double(v)*factor);                      // monte-carlo.h:59
            if (rec_samples) for (int i=0;i<D;++i) rec_samples[(k*spp+s)*D+i]=x[i];
            s1 += double(v); s2 += double(v)*double(v);
        }
        bins[k] = float(double(nb)*double(sol));                              // integrator-per-bin-parallel.h:33 ('=')
        if (rec_sum) rec_sum[k]=s1;
        if (rec_sum2) rec_sum2[k]=s2;
    }
    return 0;
}

// monte-carlo.h:39-63
extern "C" int vo_monte_carlo(const char* integrand, int dimbins, const uint64_t* res,
                   const float* rmin, const float* rmax, uint64_t samples, uint64_t seed,
                   float* bins, float* rec_samples) {
    auto F = find_finite(integrand); if (!F) return -1;
    int D = F->dim; if (dimbins<1 || dimbins>D) return -2;
    double resolution_factor = 1; for (int i=0;i<dimbins;++i) resolution_factor *= double(res[i]);
    double factor = resolution_factor*double(volume_of(D,rmin,rmax))/double(samples);   // :45
    MT19937 rng(seed);
    for (uint64_t s=0;s<samples;++s) {
        float x[8];
        for (int i=0;i<D;++i) x[i] = uniform_real(rng,rmin[i],rmax[i]);       // :50-53
        if (rec_samples) for (int i=0;i<D;++i) rec_samples[s*D+i]=x[i];
        uint64_t lin=0, prod=1; bool ok=true;
        for (int i=0;i<dimbins;++i) {
            uint64_t p = uint64_t(float(res[i])*(x[i]-rmin[i])/(rmax[i]-rmin[i]));   // :57
            if (p>=res[i]) ok=false;   // the reference would write out of bounds here (x rounded up to max); never hit in tests
            lin += p*prod; prod *= res[i];
        }
        float v = F->fn(x);
        if (ok) bins[lin] = float(double(bins[lin]) + double(v)*factor);      // :59
    }
    return 0;
}

// monte-carlo-per-bin-parallel.h:73-100 + random-sequence-ref-dis.h:11-44 + range-infinite.h:16-64
extern "C" int vo_mc_per_bin_parallel_inf(const char* integrand, int dimbins, const uint64_t* res,
                               const float* rmin, const float* rmax, int nrange,
                               uint64_t spp, uint64_t seed, float* bins,
                               double* rec_sum, double* rec_sum2,
                               uint32_t* rec_len, float* rec_elems, uint64_t rec_cap, uint64_t* rec_used) {
    auto F = find_inf(integrand); if (!F) return -1;
    if (dimbins<1 || dimbins>8) return -2;
    auto rmin_at = [&] (int i) { return i<nrange ? rmin[i] : 0.0f; };       // range-infinite.h:31-37
    auto rmax_at = [&] (int i) { return i<nrange ? rmax[i] : 1.0f; };
    float vol = 1.0f; for (int i=0;i<nrange;++i) vol *= (rmax_at(i)-rmin_at(i));        // range-infinite.h:22-23
    double factor = vol/double(spp);                                          // :77
    uint64_t nb = nbins_of(dimbins,res);
    MT19937 master(seed);
    std::vector<uint32_t> perbin_seed(nb);
    for (uint64_t k=0;k<nb;++k) perbin_seed[k] = uint32_t(master());          // :83-87
    uint64_t used = 0; int rc = 0;
    for (uint64_t k=0;k<nb;++k) {
        uint64_t pos[8]; unflatten(k,dimbins,res,pos);
        MT19937 local(perbin_seed[k]);
        int nsub = std::max(nrange,dimbins);
        std::vector<float> a(nsub), b(nsub);
        for (int i=0;i<nsub;++i) { a[i]=rmin_at(i); b[i]=rmax_at(i); }
        for (int i=0;i<dimbins;++i) {                                          // :79,93-94
            float drange = (rmax_at(i)-rmin_at(i))/float(res[i]);
            a[i] = rmin_at(i)+float(pos[i])*drange; b[i] = rmin_at(i)+float(pos[i]+1)*drange;
        }
        double s1=0, s2=0;
        for (uint64_t s=0;s<spp;++s) {
            uint32_t count = 0; int idx = 0;
            LazySeq seq;
            seq.next = [&] () -> float {                                      // random-sequence-ref-dis.h:28,32
                int i = idx++;
                float lo = i<nsub ? a[i] : 0.0f, hi = i<nsub ? b[i] : 1.0f;
                float n = uniform_real(local,0.0f,1.0f)*(hi-lo)+lo;
                ++count;
                if (rec_elems) { if (used < rec_cap) rec_elems[used] = n; else rc = -3; }
                ++used;
                return n;
            };
            float v = F->fn(seq);
            bins[k] = float(double(bins[k]) + double(v)*factor);              // :96
            if (rec_len) rec_len[k*spp+s] = count;
            s1 += double(v); s2 += double(v)*double(v);
        }
        if (rec_sum) rec_sum[k]=s1;
        if (rec_sum2) rec_sum2[k]=s2;
    }
    if (rec_used) *rec_used = used;
    return rc;
}

// integrator-per-bin-parallel.h:16-35 wrapping monte-carlo.h:65-84 (RangeInfinite) through integrate.h:115-123;
// sequences are RandomSequenceRNG (random-sequence-rng.h:11-45): a fresh mt19937(seed) per sample, element i =
// uniform_real_distribution(min_i,max_i)(rng)
extern "C" int vo_per_bin_parallel_mc_inf(const char* integrand, int dimbins, const uint64_t* res,
                               const float* rmin, const float* rmax, int nrange,
                               uint64_t spp, uint64_t seed, float* bins,
                               double* rec_sum, double* rec_sum2,
                               uint32_t* rec_len, float* rec_elems, uint64_t rec_cap, uint64_t* rec_used) {
    auto F = find_inf(integrand); if (!F) return -1;
    if (dimbins<1 || dimbins>8) return -2;
    auto rmin_at = [&] (int i) { return i<nrange ? rmin[i] : 0.0f; };
    auto rmax_at = [&] (int i) { return i<nrange ? rmax[i] : 1.0f; };
    uint64_t nb = nbins_of(dimbins,res);
    MT19937 user(seed);
    MT19937 master{uint64_t(user())};                                          // copy into the wrapper reseeds (monte-carlo.h:32-33)
    std::vector<uint32_t> perbin_seed(nb);
    for (uint64_t k=0;k<nb;++k) perbin_seed[k] = master();                     // tensor<Integrator> copies, tensor order
    uint64_t used = 0; int rc = 0;
    for (uint64_t k=0;k<nb;++k) {
        uint64_t pos[8]; unflatten(k,dimbins,res,pos);
        MT19937 binrng(perbin_seed[k]);
        int nsub = std::max(nrange,dimbins);
        std::vector<float> a(nsub), b(nsub);
        for (int i=0;i<nsub;++i) { a[i]=rmin_at(i); b[i]=rmax_at(i); }
        for (int i=0;i<dimbins;++i) {
            float drange = (rmax_at(i)-rmin_at(i))/float(res[i]);
            a[i] = rmin_at(i)+float(pos[i])*drange; b[i] = rmin_at(i)+float(pos[i]+1)*drange;
        }
        float vol = 1.0f; for (int i=0;i<nsub;++i) vol *= (b[i]-a[i]);          // RangeInfinite::_volume of the bin sub-range
        double factor = 1.0*double(vol)/double(spp);                           // monte-carlo.h:70-72, one bin
        float sol = 0.0f;
        double s1=0, s2=0;
        for (uint64_t s=0;s<spp;++s) {
            MT19937 seqrng{uint64_t(uint32_t(binrng()))};                       // random_sequence(range, unsigned seed) -> mt19937(seed)  (:75)
            // (:76-80 walks a separate copy of the stream to find the bin position — always bin 0 here — no effect on f's stream)
            uint32_t count = 0; int idx = 0;
            LazySeq seq;
            seq.next = [&] () -> float {
                int i = idx++;
                float lo = i<nsub ? a[i] : 0.0f, hi = i<nsub ? b[i] : 1.0f;
                float n = uniform_real(seqrng,lo,hi);                           // random-sequence-rng.h:30,35
                ++count;
                if (rec_elems) { if (used < rec_cap) rec_elems[used] = n; else rc = -3; }
                ++used;
                return n;
            };
            float v = F->fn(seq);
            sol = float(double(sol) + double(v)*factor);                       // :81
            if (rec_len) rec_len[k*spp+s] = count;
            s1 += double(v); s2 += double(v)*double(v);
        }
        bins[k] = float(double(nb)*double(sol));                               // integrator-per-bin-parallel.h:33
        if (rec_sum) rec_sum[k]=s1;
        if (rec_sum2) rec_sum2[k]=s2;
    }
    if (rec_used) *rec_used = used;
    return rc;
}

// monte-carlo.h:65-84 — global sampler over RangeInfinite, bin from the first dimbins sequence elements
extern "C" int vo_monte_carlo_inf(const char* integrand, int dimbins, const uint64_t* res,
                       const float* rmin, const float* rmax, int nrange, uint64_t samples, uint64_t seed, float* bins) {
    auto F = find_inf(integrand); if (!F) return -1;
    if (dimbins<1 || dimbins>8) return -2;
    auto rmin_at = [&] (int i) { return i<nrange ? rmin[i] : 0.0f; };
    auto rmax_at = [&] (int i) { return i<nrange ? rmax[i] : 1.0f; };
    double resolution_factor = 1; for (int i=0;i<dimbins;++i) resolution_factor *= double(res[i]);
    float vol = 1.0f; for (int i=0;i<nrange;++i) vol *= (rmax_at(i)-rmin_at(i));
    double factor = resolution_factor*double(vol)/double(samples);             // :70-72
    MT19937 rng(seed);
    for (uint64_t s=0;s<samples;++s) {
        uint32_t sd = uint32_t(rng());                                          // :75
        MT19937 posrng{uint64_t(sd)};                                           // :76-80: a copy of the stream gives the bin position
        uint64_t lin=0, prod=1; bool ok=true;
        for (int i=0;i<dimbins;++i) {
            float e = uniform_real(posrng,rmin_at(i),rmax_at(i));
            uint64_t p = uint64_t(float(res[i])*(e-rmin_at(i))/(rmax_at(i)-rmin_at(i)));
            if (p>=res[i]) ok=false;
            lin += p*prod; prod *= res[i];
        }
        MT19937 seqrng{uint64_t(sd)};
        int idx = 0;
        LazySeq seq;
        seq.next = [&] () -> float { int i = idx++; return uniform_real(seqrng,rmin_at(i),rmax_at(i)); };
        float v = F->fn(seq);
        if (ok) bins[lin] = float(double(bins[lin]) + double(v)*factor);       // :81
    }
    return 0;
}

// =========================================================================================================
// Newton-Cotes rules (rules.h), nested pairs (nested.h), error metrics / heuristics (error-*.h)
// =========================================================================================================
namespace {

enum Rule { TRAPEZOIDAL=2, SIMPSON=3, BOOLE=5 };   // value == samples per dimension
// Everything below is a template over the scalar type T = Float = value_type of the reference (float or double): the
// float(...) casts of the fp32 restatement become T(...), i.e. no-ops when the reference itself computes in double.
template<typename T> using FiniteFnT = T (*)(const T*);
template<typename T> inline T volume_of_t(int d, const T* a, const T* b) { T v=T(1); for (int i=0;i<d;++i) v*=(b[i]-a[i]); return v; }
template<typename T> inline void bin_box_t(int d, int db, const T* rmin, const T* rmax, const uint64_t* res, const uint64_t* pos, T* a, T* b) {
    for (int i=0;i<d;++i) { a[i]=rmin[i]; b[i]=rmax[i]; }
    for (int i=0;i<db;++i) {
        T drange = (rmax[i]-rmin[i])/T(res[i]);
        a[i] = rmin[i] + T(pos[i])*drange;
        b[i] = rmin[i] + T(pos[i]+1)*drange;
    }
}


// rules.h:14 / :64 / :256 — literals are double, result rounded to T at return
template<typename T> inline T rule_apply(int S, const T* p) {
    switch (S) {
        case 2: return T((p[0]+p[1])/2.0);                                               // T add, double divide
        case 3: return T((double(p[0])+4.0*double(p[1])+double(p[2]))/6.0);
        default: return T((7.0*double(p[0])+32.0*double(p[1])+12.0*double(p[2])+32.0*double(p[3])+7.0*double(p[4]))/90.0);
    }
}
// rules.h:27-32 / :77-83 / :272-280 — integer literals: T arithmetic, left to right
template<typename T> inline void rule_coefficients(int S, const T* p, T* c) {
    switch (S) {
        case 2: c[0]=p[0]; c[1]=p[1]-p[0]; break;
        case 3: c[0]=p[0]; c[1]=-3*p[0]+4*p[1]-p[2]; c[2]=2*p[0]-4*p[1]+2*p[2]; break;
        default:
            c[0]=p[0];
            c[1]=-25*p[0]/3+ 16*p[1] - 12*p[2] +16*p[3]/3 - p[4];
            c[2]=70*p[0]/3 -208*p[1]/3 + 76*p[2] -112*p[3]/3 + 22*p[4]/3;
            c[3]=-80*p[0]/3 + 96*p[1] - 128*p[2] + 224*p[3]/3 - 16*p[4];
            c[4]=32*p[0]/3 - 128*p[1]/3 + 64*p[2] - 128*p[3]/3 + 32*p[4]/3;
    }
}
// rules.h:35-38 / :86-89 / :283-286 — Horner in T
template<typename T> inline T rule_at(int S, T t, const T* p) {
    T c[5]; rule_coefficients(S,p,c);
    switch (S) {
        case 2: return c[1]*t + c[0];
        case 3: return (c[2]*t + c[1])*t + c[0];
        default: return (((c[4]*t + c[3])*t + c[2])*t + c[1])*t + c[0];
    }
}
// rules.h:41-44 / :97-100 / :289-293 — antiderivative difference; c*b is T, the /k.0 promotes to double
template<typename T> inline T rule_subrange(int S, T a, T b, const T* p) {
    T c[5]; rule_coefficients(S,p,c);
    switch (S) {
        case 2: return T((c[1]*b/2.0 + c[0])*b - (c[1]*a/2.0 + c[0])*a);
        case 3: return T(((c[2]*b/3.0 + c[1]/2.0)*b + c[0])*b - ((c[2]*a/3.0 + c[1]/2.0)*a + c[0])*a);
        default: return T(((((c[4]*b/5.0 + c[3]/4.0)*b + c[2]/3.0)*b + c[1]/2.0)*b + c[0])*b -
                              ((((c[4]*a/5.0 + c[3]/4.0)*a + c[2]/3.0)*a + c[1]/2.0)*a + c[0])*a);
    }
}
// nested.h:17-23
template<typename T> inline T nested_low(int SH, int SL, const T* p) {
    T plow[5];
    for (int i=0;i<SL;++i) plow[i] = p[i*(SH-1)/(SL-1)];
    return rule_apply(SL,plow);
}
// error-metric.h:10-13 / :30-37
template<typename T> inline T metric_absolute(T a, T b) { return std::abs(b-a); }
template<typename T> inline T metric_relative(T a, T b) {
    const double min_val = 1.e-37;
    if (std::max(std::abs(a),std::abs(b)) < min_val) return std::abs(b-a);
    else return std::abs(b-a)/std::max(std::abs(a),std::abs(b));
}

inline uint64_t ipow(int s, int d) { uint64_t r=1; for (int i=0;i<d;++i) r*=uint64_t(s); return r; }

// A region = box + S^D samples, dim-0-fastest (region.h:23-68, multiarray.h:20-25)
template<typename T> struct RegionT {
    std::vector<T> rmin, rmax, data;
    T volume;                     // Range::_volume, T product
    T err = T(0); uint32_t errdim = 0;
    double key = 0.0;             // what the heap orders by: double(err) for the Float-keyed heuristics, error_heuristic_mixed's own double key
};

// fold a multiarray along dimension `dim` with a per-line functor; shape in/out are S^nd / S^(nd-1), dim-0-fastest
// (fold.h:24-35: result index = base index with `dim` removed)
template<typename T, typename LineFn>
std::vector<T> fold_dim(const std::vector<T>& in, int S, int nd, int dim, LineFn&& fn) {
    uint64_t inner = ipow(S,dim), outer = ipow(S,nd-1-dim);
    std::vector<T> out(inner*outer);
    T line[5];
    for (uint64_t o=0;o<outer;++o) for (uint64_t i=0;i<inner;++i) {
        for (int e=0;e<S;++e) line[e] = in[i + uint64_t(e)*inner + o*inner*uint64_t(S)];
        out[i + o*inner] = fn(line);
    }
    return out;
}
// fold_all(q): fold dimension 0 repeatedly (fold.h:87-108) — the lowest remaining dim is always innermost
template<typename T> inline T fold_all_rule(std::vector<T> v, int S, int nd) {
    while (nd>0) { v = fold_dim(v,S,nd,0,[&] (const T* l) { return rule_apply(S,l); }); --nd; }
    return v[0];
}

// region.h:40-46 (f_in_range) + fill.h:45-72: normalised coords double(i)/double(S-1), mapped in double, rounded to T
template<typename T> inline void region_point(const RegionT<T>& r, int D, const double* p, T* x) {
    for (int i=0;i<D;++i) x[i] = T(p[i]*double(r.rmax[i]-r.rmin[i]) + double(r.rmin[i]));
}
template<typename T, typename Fn> RegionT<T> make_region(Fn&& f, int S, int D, const T* a, const T* b) {
    RegionT<T> r; r.rmin.assign(a,a+D); r.rmax.assign(b,b+D); r.volume = volume_of_t(D,a,b);
    uint64_t n = ipow(S,D); r.data.resize(n);
    for (uint64_t k=0;k<n;++k) {
        double p[8]; T x[8]; uint64_t t=k;
        for (int d=0;d<D;++d) { p[d] = double(t%uint64_t(S))/double(S-1); t/=uint64_t(S); }
        region_point(r,D,p,x);
        r.data[k] = f(x);
    }
    return r;
}

// region.h:345-359 + split.h:13-49, parts == 2.  Children reuse the parent's samples at even positions along
// `dim` and evaluate f at the odd ones with coordinates derived from the PARENT range (v = i/(2(S-1))).
template<typename T, typename Fn> void split_region(Fn&& f, int S, int D, const RegionT<T>& r, int dim, RegionT<T> out[2]) {
    uint64_t n = ipow(S,D), inner = ipow(S,dim);
    int full = 2*(S-1)+1;
    for (int c=0;c<2;++c) { out[c].rmin=r.rmin; out[c].rmax=r.rmax; out[c].data.assign(n,T(0)); }
    // slab(i) of the conceptual (2S-1)-wide array: from parent if i even, fresh evaluations if odd
    for (int i=0;i<full;++i) {
        for (uint64_t k=0;k<n/uint64_t(S);++k) {           // index over the other D-1 dims, dim-0-fastest
            uint64_t lo = k%inner, hi = k/inner;            // position below / above `dim`
            T v;
            if (i%2==0) v = r.data[lo + uint64_t(i/2)*inner + hi*inner*uint64_t(S)];
            else {
                double p[8]; T x[8]; uint64_t t=k;
                for (int d=0;d<D;++d) {
                    if (d==dim) p[d] = double(i)/double(full-1);
                    else { p[d] = double(t%uint64_t(S))/double(S-1); t/=uint64_t(S); }
                }
                region_point(r,D,p,x);
                v = f(x);
            }
            if (i<=S-1) out[0].data[lo + uint64_t(i)*inner + hi*inner*uint64_t(S)] = v;
            if (i>=S-1) out[1].data[lo + uint64_t(i-(S-1))*inner + hi*inner*uint64_t(S)] = v;
        }
    }
    T d = (r.rmax[dim]-r.rmin[dim])/T(2);                              // region.h:351
    T mid = r.rmin[dim] + d*T(1);                                      // :353 (i+1 == 1)
    out[0].rmax[dim] = mid; out[1].rmin[dim] = mid;                            // :353-355 ; last child keeps parent max
    for (int c=0;c<2;++c) out[c].volume = volume_of_t(D,out[c].rmin.data(),out[c].rmax.data());
}

// region.h:387-393: volume * fold(error along dim).fold_all(high rule)
template<typename T> T region_error(const RegionT<T>& r, int SH, int SL, int D, int dim, bool relative) {
    auto e = fold_dim(r.data,SH,D,dim,[&] (const T* l) {
        T h = rule_apply(SH,l), lo = nested_low(SH,SL,l);
        return relative ? metric_relative(h,lo) : metric_absolute(h,lo);       // nested.h:31-33
    });
    return r.volume*fold_all_rule(e,SH,D-1);
}
// region.h:420-424 Region::error(): volume * (fold_all(high) - fold_all(low))
template<typename T> T region_error_total(const RegionT<T>& r, int SH, int SL, int D) {
    std::vector<T> v = r.data; int nd = D;
    while (nd>0) { v = fold_dim(v,SH,nd,0,[&] (const T* l) { return nested_low(SH,SL,l); }); --nd; }
    return r.volume*(fold_all_rule(r.data,SH,D) - v[0]);
}
// error-heuristic.h:15-18 -> region.h:401-411 (first maximal dim, strict >)
template<typename T> void heuristic_default(RegionT<T>& r, int SH, int SL, int D, bool relative) {
    T max_err = 0; uint32_t max_dim = 0;
    for (int d=0; d<D; ++d) { T err = region_error(r,SH,SL,D,d,relative); if (err>max_err) { max_err=err; max_dim=uint32_t(d); } }
    r.err = max_err; r.errdim = max_dim; r.key = double(max_err);
}
// error-heuristic.h:29-46 (last maximal dim, >=)
template<typename T> void heuristic_size(RegionT<T>& r, int SH, int SL, int D, bool relative, double size_weight) {
    const double min_size = 1.e-37;
    T max_err = region_error(r,SH,SL,D,0,relative);
    T w0 = r.rmax[0]-r.rmin[0];
    if (w0<min_size || std::isnan(w0)) max_err = 0;
    else max_err = T(double(max_err) + size_weight*double(std::abs(r.rmax[0]-r.rmin[0])));
    uint32_t max_dim = 0;
    T err = max_err;
    for (int d=1; d<D; ++d) {
        err = T(double(region_error(r,SH,SL,D,d,relative)) + size_weight*double(std::abs(r.rmax[d]-r.rmin[d])));
        if ((r.rmax[d]-r.rmin[d])<min_size) err = 0;
        if (err>=max_err) { max_err=err; max_dim=uint32_t(d); }
    }
    r.err = max_err; r.errdim = max_dim; r.key = double(max_err);
}
// error-heuristic.h:49-98 error_heuristic_mixed: two metrics (bins: d = 0 and d < dimension; rest: beyond), a DOUBLE key
struct MixedArgs { int dimension = 2; double bins_weight = 1.0, size_threshold_bins = 1.0/1024.0, size_threshold_rest = 1.0/16.0, error_increase_factor = 1.e4; };
MixedArgs g_mixed;
template<typename T> void heuristic_mixed(RegionT<T>& r, int SH, int SL, int D, bool rel_bins, bool rel_rest, double size_weight, const MixedArgs& m) {
    double size_bins = 1.0, size_rest = 1.0;
    for (int d=0; d<std::min(m.dimension,D); ++d) size_bins *= std::abs(r.rmax[d]-r.rmin[d]);                      // :73-74
    for (int d=m.dimension; d<D; ++d) size_rest *= std::abs(r.rmax[d]-r.rmin[d]);                                  // :75-76
    double add_bins = m.error_increase_factor, add_rest = m.error_increase_factor;                                 // :78-83
    if (size_bins<m.size_threshold_bins) add_bins = 0.0;
    if (size_rest<m.size_threshold_rest) add_rest = 0.0;
    if (std::isnan(size_bins)) add_bins = 0.0;
    if (std::isnan(size_rest)) add_rest = 0.0;
    double max_err = region_error(r,SH,SL,D,0,rel_bins)*m.bins_weight + add_bins + size_weight*(r.rmax[0]-r.rmin[0]);    // :85-86
    uint32_t max_dim = 0;
    double err = max_err;
    for (int d=1; d<D; ++d) {                                                                                      // :90-97
        if (d<m.dimension) err = region_error(r,SH,SL,D,d,rel_bins)*m.bins_weight + add_bins + size_weight*(r.rmax[d]-r.rmin[d]);
        else err = region_error(r,SH,SL,D,d,rel_rest) + add_rest + size_weight*(r.rmax[d]-r.rmin[d]);
        if (err>=max_err) { max_err=err; max_dim=uint32_t(d); }
    }
    r.err = T(max_err); r.errdim = max_dim; r.key = max_err;
}

struct Heuristic { bool size; bool relative; double size_weight; bool mixed = false; bool relative_rest = false; MixedArgs margs; };
bool parse_heuristic(const char* h, double sw, Heuristic& out) {
    out.size_weight = sw; out.mixed = false;
    if (!std::strncmp(h,"mixed_",6)) {                 // mixed_<bins metric>_<rest metric>; the other arguments come from vo_set_mixed
        const char* rest = std::strchr(h+6,'_'); if (!rest) return false;
        std::string mb(h+6, rest), mr(rest+1);
        if ((mb!="absolute" && mb!="relative") || (mr!="absolute" && mr!="relative")) return false;
        out.mixed = true; out.size = false; out.relative = mb=="relative"; out.relative_rest = mr=="relative"; out.margs = g_mixed;
        return true;
    }
    if (!std::strcmp(h,"default_absolute")) { out.size=false; out.relative=false; return true; }
    if (!std::strcmp(h,"default_relative")) { out.size=false; out.relative=true;  return true; }
    if (!std::strcmp(h,"size_absolute"))    { out.size=true;  out.relative=false; return true; }
    if (!std::strcmp(h,"size_relative"))    { out.size=true;  out.relative=true;  return true; }
    return false;
}
template<typename T> void apply_heuristic(RegionT<T>& r, int SH, int SL, int D, const Heuristic& h) {
    if (h.mixed) { heuristic_mixed(r,SH,SL,D,h.relative,h.relative_rest,h.size_weight,h.margs); return; }
    if (h.size) heuristic_size(r,SH,SL,D,h.relative,h.size_weight); else heuristic_default(r,SH,SL,D,h.relative);
}

// libstdc++ bits/stl_heap.h:135-148 (__push_heap) with comparator a.err < b.err
template<typename T> void heap_push(std::vector<RegionT<T>>& h) {
    std::size_t hole = h.size()-1; RegionT<T> value = std::move(h[hole]);
    std::size_t parent = (hole-1)/2;
    while (hole>0 && h[parent].key < value.key) { h[hole] = std::move(h[parent]); hole = parent; parent = (hole-1)/2; }
    h[hole] = std::move(value);
}
// libstdc++ bits/stl_heap.h:254-267 (pop_heap -> __pop_heap) + :224-250 (__adjust_heap); caller pops the back
template<typename T> void heap_pop(std::vector<RegionT<T>>& h) {
    if (h.size()<2) return;
    std::size_t last = h.size()-1;
    RegionT<T> value = std::move(h[last]); h[last] = std::move(h[0]);
    std::size_t len = last, hole = 0, child = 0;
    while (child < (len-1)/2) {
        child = 2*(child+1);
        if (h[child].key < h[child-1].key) --child;
        h[hole] = std::move(h[child]); hole = child;
    }
    if ((len&1)==0 && child==(len-2)/2) { child = 2*(child+1); h[hole] = std::move(h[child-1]); hole = child-1; }
    // __push_heap(first, hole, top=0, value)
    std::size_t parent = (hole-1)/2;
    while (hole>0 && h[parent].key < value.key) { h[hole] = std::move(h[parent]); hole = parent; parent = (hole-1)/2; }
    h[hole] = std::move(value);
}

// regions-generator-adaptive-heap.h:18-45
template<typename T, typename Fn> std::vector<RegionT<T>> generate_adaptive(Fn&& f, int SH, int SL, int D, const T* rmin, const T* rmax,
                                      const Heuristic& h, uint64_t iterations) {
    std::vector<RegionT<T>> heap; heap.reserve(iterations+1);
    heap.push_back(make_region<T>(f,SH,D,rmin,rmax));
    apply_heuristic(heap[0],SH,SL,D,h);
    for (uint64_t i=0;i<iterations;++i) {
        RegionT<T> r = heap.front();                                               // :33
        RegionT<T> sub[2]; split_region<T>(f,SH,D,r,int(r.errdim),sub);               // :34
        heap_pop(heap); heap.pop_back();                                       // :35
        for (int c=0;c<2;++c) {                                                // :36-40
            apply_heuristic(sub[c],SH,SL,D,h);
            heap.push_back(std::move(sub[c])); heap_push(heap);
        }
    }
    return heap;
}

// range.h:45-53
template<typename T> inline T pos_in_range1(T lo, T hi, T p) { return (lo>=hi) ? lo : (p-lo)/(hi-lo); }

// region.h:141-169 (integral_subrange -> sub_last): fold subrange(a_d,b_d) over dim D-1, ..., 0; times volume
template<typename T> T region_integral_subrange(const RegionT<T>& r, int S, int D, const T* a, const T* b) {
    std::vector<T> v = r.data;
    for (int d=D-1; d>=0; --d) {
        T na = pos_in_range1(r.rmin[d],r.rmax[d],a[d]), nb = pos_in_range1(r.rmin[d],r.rmax[d],b[d]);
        v = fold_dim(v,S,d+1,d,[&] (const T* l) { return rule_subrange(S,na,nb,l); });
    }
    return r.volume*v[0];
}
// Simpson::pdf_points / pdf_integral_subrange (rules.h:104-154): |p| shifted so that its parabola never dips below zero, integrated over [t0,t1]
template<typename T> inline T simpson_pdf_integral_subrange(T t0, T t1, const T* p) {
    T q[3]; for (int i=0;i<3;++i) q[i] = std::abs(p[i]);                        // NormDefault (norm.h:12)
    T cs[5]; rule_coefficients(3,q,cs);
    T ymin = 0;
    if (cs[2] > 0) {
        T tmin = T(-cs[1]/(2.0*cs[2]));                                        // rules.h:123 (double literal)
        if ((tmin > 0) && (tmin < 1)) { T ytmin = (cs[2]*tmin + cs[1])*tmin + cs[0]; if (ytmin < ymin) ymin = ytmin; }
    }
    for (int i=0;i<3;++i) q[i] -= ymin;
    return rule_subrange(3,t0,t1,q);
}
// region.h:277-302 (pdf_integral_subrange -> pdf_sub): fold pdf_integral_subrange(a_d,b_d) over dim D-1, ..., 0; times volume
template<typename T> T region_pdf_integral_subrange(const RegionT<T>& r, int D, const T* a, const T* b) {
    std::vector<T> v = r.data;
    for (int d=D-1; d>=0; --d) {
        T na = pos_in_range1(r.rmin[d],r.rmax[d],a[d]), nb = pos_in_range1(r.rmin[d],r.rmax[d],b[d]);
        v = fold_dim(v,3,d+1,d,[&] (const T* l) { return simpson_pdf_integral_subrange(na,nb,l); });
    }
    return r.volume*v[0];
}
// region.h:86-112 (approximation_at -> app_at): fold at(pos_d) over dim 0, 1, ..., D-1; times volume_from(D) == 1
template<typename T> T region_approximation_at(const RegionT<T>& r, int S, int D, const T* pos) {
    std::vector<T> v = r.data;
    for (int d=0; d<D; ++d) {
        T t = pos_in_range1(r.rmin[d],r.rmax[d],pos[d]);
        v = fold_dim(v,S,D-d,0,[&] (const T* l) { return rule_at(S,t,l); });
    }
    return v[0]*T(1);
}

// region.h:454-463
template<typename T> void pixels_in_region(const RegionT<T>& r, int db, const uint64_t* res, const T* rmin, const T* rmax, uint64_t* start, uint64_t* end) {
    for (int i=0;i<db;++i) {
        start[i] = std::max(uint64_t(0), uint64_t(T(res[i])*(r.rmin[i]-rmin[i])/(rmax[i]-rmin[i])));
        end[i]   = std::max(start[i]+1, std::min(res[i], uint64_t(T(0.99f) + (T(res[i])*(r.rmax[i]-rmin[i])/(rmax[i]-rmin[i])))));
    }
}
// range.h:92-101 ; false when the intersection is empty (range.h:70-75)
template<typename T> bool intersect(int D, const T* a1, const T* b1, const T* a2, const T* b2, T* a, T* b) {
    bool empty = false;
    for (int d=0;d<D;++d) { a[d] = std::max(a1[d],a2[d]); b[d] = std::max(a[d], std::min(b1[d],b2[d])); if (a[d]>=b[d]) empty = true; }
    return !empty;
}

// regions-integrator-sequential.h:38-58
template<typename T> void integrate_regions_sequential(const std::vector<RegionT<T>>& regions, int S, int D, int db, const uint64_t* res,
                                  const T* rmin, const T* rmax, T* bins) {
    uint64_t factor = nbins_of(db,res);
    for (const RegionT<T>& r : regions) {
        uint64_t st[8], en[8]; pixels_in_region(r,db,res,rmin,rmax,st,en);
        uint64_t pos[8]; for (int i=0;i<db;++i) pos[i]=st[i];
        while (true) {
            T ba[8], bb[8], ia[8], ib[8];
            bin_box_t(D,db,rmin,rmax,res,pos,ba,bb);
            if (intersect(D,ba,bb,r.rmin.data(),r.rmax.data(),ia,ib)) {
                uint64_t lin=0, prod=1; for (int i=0;i<db;++i) { lin+=pos[i]*prod; prod*=res[i]; }
                bins[lin] = T(double(bins[lin]) + double(factor)*double(region_integral_subrange(r,S,D,ia,ib)));   // :54
            }
            int d=0; for (; d<db; ++d) { if (++pos[d] >= en[d]) pos[d]=st[d]; else break; }
            if (d==db) break;
        }
    }
}

// ---- Steps<Q,N> composite rules (rules.h:321-388): (Q-1)*N+1 samples per dimension, N pieces of rule Q ---------------------------
template<typename T> inline T steps_subrange(int Q, int N, T a, T b, const T* p) {                    // rules.h:359-384
    std::size_t ia = std::min(std::size_t(a*N), std::size_t(N-1));
    std::size_t ib = std::min(std::size_t(b*N), std::size_t(N-1));
    T a_local = T(a*N) - T(ia);
    T b_local = T(b*N) - T(ib);
    if (ia == ib) return rule_subrange(Q,a_local,b_local,p+ia*(Q-1))/N;
    T sol = rule_subrange(Q,a_local,T(1),p+ia*(Q-1))/N;
    for (std::size_t i = ia+1; i<ib; ++i) sol += rule_apply(Q,p+i*(Q-1))/N;
    sol += rule_subrange(Q,T(0),b_local,p+ib*(Q-1))/N;
    return sol;
}
// Region::integral_subrange with a Steps rule (region.h:141-169: fold subrange over dim D-1, ..., 0; times the region's volume)
template<typename T> T region_integral_subrange_steps(const RegionT<T>& r, int Q, int N, int D, const T* a, const T* b) {
    const int S = (Q-1)*N+1;
    std::vector<T> v = r.data;
    std::vector<T> line(S);
    for (int d=D-1; d>=0; --d) {
        T na = pos_in_range1(r.rmin[d],r.rmax[d],a[d]), nb = pos_in_range1(r.rmin[d],r.rmax[d],b[d]);
        uint64_t inner = ipow(S,d);                      // dims below d; d is the last remaining dimension
        std::vector<T> out(inner);
        for (uint64_t i=0;i<inner;++i) {
            for (int e=0;e<S;++e) line[e] = v[i + uint64_t(e)*inner];
            out[i] = steps_subrange(Q,N,na,nb,line.data());
        }
        v.swap(out);
    }
    return r.volume*v[0];
}
template<typename T> void integrate_steps_region(const RegionT<T>& r, int Q, int N, int D, int db, const uint64_t* res, const T* rmin, const T* rmax, T* bins) {
    uint64_t factor = nbins_of(db,res);
    uint64_t st[8], en[8]; pixels_in_region(r,db,res,rmin,rmax,st,en);
    uint64_t pos[8]; for (int i=0;i<db;++i) pos[i]=st[i];
    while (true) {
        T ba[8], bb[8], ia[8], ib[8];
        bin_box_t(D,db,rmin,rmax,res,pos,ba,bb);
        if (intersect(D,ba,bb,r.rmin.data(),r.rmax.data(),ia,ib)) {
            uint64_t lin=0, prod=1; for (int i=0;i<db;++i) { lin+=pos[i]*prod; prod*=res[i]; }
            bins[lin] = T(double(bins[lin]) + double(factor)*double(region_integral_subrange_steps(r,Q,N,D,ia,ib)));   // regions-integrator-sequential.h:54
        }
        int d=0; for (; d<db; ++d) { if (++pos[d] >= en[d]) pos[d]=st[d]; else break; }
        if (d==db) break;
    }
}
// "steps<N>_<rule>" -> (Q, N)
inline bool parse_steps(const char* rule, int* Q, int* N) {
    int n = 0; char base[32] = {0};
    if (std::sscanf(rule,"steps%d_%31s",&n,base) != 2 || n < 1) return false;
    int q = !std::strcmp(base,"trapezoidal") ? 2 : !std::strcmp(base,"simpson") ? 3 : !std::strcmp(base,"boole") ? 5 : 0;
    if (!q) return false;
    *Q = q; *N = n; return true;
}

template<typename T> void export_regions(const std::vector<RegionT<T>>& regions, int D,
                    T* reg_min, T* reg_max, T* reg_err, uint32_t* reg_dim, T* reg_data) {
    uint64_t n = 0;
    for (const RegionT<T>& r : regions) {
        if (reg_min) std::copy(r.rmin.begin(), r.rmin.end(), reg_min+n*uint64_t(D));
        if (reg_max) std::copy(r.rmax.begin(), r.rmax.end(), reg_max+n*uint64_t(D));
        if (reg_err) reg_err[n] = r.err;
        if (reg_dim) reg_dim[n] = r.errdim;
        if (reg_data) std::copy(r.data.begin(), r.data.end(), reg_data+n*r.data.size());
        ++n;
    }
}

} // namespace

// newton-cotes.h:11-14 = regions-generator-single.h:12-20 + regions-integrator-sequential.h:38-58
extern "C" int vo_newton_cotes(const char* integrand, const char* rule, int dimbins, const uint64_t* res,
                    const float* rmin, const float* rmax, float* bins) {
    auto F = find_finite(integrand); if (!F) return -1;
    int D = F->dim; if (dimbins<1 || dimbins>D) return -2;
    int SQ, SN;
    if (parse_steps(rule,&SQ,&SN)) {                                           // integrator_newton_cotes(steps<N>(rule)), rules.h:321-388
        RegionT<float> r = make_region<float>(F->fn,(SQ-1)*SN+1,D,rmin,rmax);
        integrate_steps_region<float>(r,SQ,SN,D,dimbins,res,rmin,rmax,bins);
        return 0;
    }
    int S = !std::strcmp(rule,"trapezoidal") ? 2 : !std::strcmp(rule,"simpson") ? 3 : !std::strcmp(rule,"boole") ? 5 : 0;
    if (!S) return -2;
    std::vector<RegionT<float>> regions; regions.push_back(make_region<float>(F->fn,S,D,rmin,rmax));
    integrate_regions_sequential<float>(regions,S,D,dimbins,res,rmin,rmax,bins);
    return 0;
}

extern "C" void vo_set_mixed(int dimension, double bins_weight, double size_threshold_bins, double size_threshold_rest, double error_increase_factor) {
    g_mixed.dimension = dimension; g_mixed.bins_weight = bins_weight; g_mixed.size_threshold_bins = size_threshold_bins;
    g_mixed.size_threshold_rest = size_threshold_rest; g_mixed.error_increase_factor = error_increase_factor;
}

extern "C" int vo_adaptive_iterations(const char* integrand, const char* rule, const char* heuristic, double size_weight,
                           uint64_t iterations, int dimbins, const uint64_t* res,
                           const float* rmin, const float* rmax, float* bins,
                           float* reg_min, float* reg_max, float* reg_err, uint32_t* reg_dim, float* reg_data) {
    auto F = find_finite(integrand); if (!F) return -1;
    int D = F->dim; if (dimbins<1 || dimbins>D) return -2;
    int SH, SL;
    if (!std::strcmp(rule,"simpson_trapezoidal")) { SH=3; SL=2; }
    else if (!std::strcmp(rule,"boole_simpson")) { SH=5; SL=3; }
    else return -2;
    Heuristic h; if (!parse_heuristic(heuristic,size_weight,h)) return -2;
    auto regions = generate_adaptive<float>(F->fn,SH,SL,D,rmin,rmax,h,iterations);
    export_regions<float>(regions,D,reg_min,reg_max,reg_err,reg_dim,reg_data);
    if (bins) integrate_regions_sequential<float>(regions,SH,D,dimbins,res,rmin,rmax,bins);
    return 0;
}

// integrator-adaptive-tolerance.h:15-39
namespace {
struct ToleranceRun {
    FiniteFn f; int SH, SL, D, dimbins; const uint64_t* res; const float* rmin; const float* rmax; Heuristic h; float tolerance;
    float* bins; uint64_t nleaves = 0; uint64_t reg_cap; float *reg_min, *reg_max, *reg_err; uint32_t* reg_dim; float* reg_data;
    bool overflow = false, too_deep = false;
    void visit(RegionT<float>& r, int depth) {
        apply_heuristic(r,SH,SL,D,h);                                           // :18
        if (r.err < tolerance) {                                                // :19
            std::vector<RegionT<float>> one(1,r);
            integrate_regions_sequential<float>(one,SH,D,dimbins,res,rmin,rmax,bins);      // :21
            if (nleaves < reg_cap) {
                uint64_t n = nleaves;
                if (reg_min) std::copy(r.rmin.begin(), r.rmin.end(), reg_min+n*uint64_t(D));
                if (reg_max) std::copy(r.rmax.begin(), r.rmax.end(), reg_max+n*uint64_t(D));
                if (reg_err) reg_err[n] = r.err;
                if (reg_dim) reg_dim[n] = r.errdim;
                if (reg_data) std::copy(r.data.begin(), r.data.end(), reg_data+n*r.data.size());
            } else if (reg_min || reg_max || reg_err || reg_dim || reg_data) overflow = true;
            ++nleaves;
        } else {
            if (depth >= 128) { too_deep = true; return; }
            RegionT<float> sub[2]; split_region<float>(f,SH,D,r,int(r.errdim),sub);          // :25
            for (int c=0;c<2 && !too_deep;++c) visit(sub[c],depth+1);                        // :26
        }
    }
};
}
extern "C" int vo_adaptive_tolerance(const char* integrand, const char* rule, const char* heuristic, double size_weight, float tolerance,
                          int dimbins, const uint64_t* res, const float* rmin, const float* rmax, float* bins, uint64_t* nleaves,
                          uint64_t reg_cap, float* reg_min, float* reg_max, float* reg_err, uint32_t* reg_dim, float* reg_data) {
    auto F = find_finite(integrand); if (!F) return -1;
    int D = F->dim; if (dimbins<1 || dimbins>D) return -2;
    ToleranceRun run;
    if (!std::strcmp(rule,"simpson_trapezoidal")) { run.SH=3; run.SL=2; }
    else if (!std::strcmp(rule,"boole_simpson")) { run.SH=5; run.SL=3; }
    else return -2;
    if (!parse_heuristic(heuristic,size_weight,run.h)) return -2;
    run.f=F->fn; run.D=D; run.dimbins=dimbins; run.res=res; run.rmin=rmin; run.rmax=rmax; run.tolerance=tolerance; run.bins=bins;
    run.reg_cap=reg_cap; run.reg_min=reg_min; run.reg_max=reg_max; run.reg_err=reg_err; run.reg_dim=reg_dim; run.reg_data=reg_data;
    RegionT<float> root = make_region<float>(F->fn,run.SH,D,rmin,rmax);       // :37
    run.visit(root,0);
    if (nleaves) *nleaves = run.nleaves;
    return run.too_deep ? -4 : run.overflow ? -3 : 0;
}

// ---- double precision (Range<double,DIM>): the same templates with T = double -----------------------------------------------
namespace {
template<typename F> double call_finite_d(const double* x) {
    std::array<double,F::dim> a; for (int i=0;i<F::dim;++i) a[i]=x[i];
    return F()(a);
}
struct FiniteIntegrandD { const char* name; int dim; FiniteFnT<double> fn; };
const FiniteIntegrandD FINITE_D[] = {
    {"x2y2",2,call_finite_d<vo::X2Y2T<double>>}, {"ind2",2,call_finite_d<vo::Ind2T<double>>}, {"cubic1",1,call_finite_d<vo::Cubic1T<double>>},
    {"poly3",3,call_finite_d<vo::Poly3T<double>>}, {"smooth_edge2",2,call_finite_d<vo::SmoothEdge2T<double>>},
    {"shade4_16",4,call_finite_d<vo::Shade4T<double,16>>},
};
const FiniteIntegrandD* find_finite_d(const char* n) { for (auto& f : FINITE_D) if (!std::strcmp(f.name,n)) return &f; return nullptr; }
}

extern "C" int vo_newton_cotes_f64(const char* integrand, const char* rule, int dimbins, const uint64_t* res,
                        const double* rmin, const double* rmax, double* bins) {
    auto F = find_finite_d(integrand); if (!F) return -1;
    int D = F->dim; if (dimbins<1 || dimbins>D) return -2;
    int S = !std::strcmp(rule,"trapezoidal") ? 2 : !std::strcmp(rule,"simpson") ? 3 : !std::strcmp(rule,"boole") ? 5 : 0;
    if (!S) return -2;
    std::vector<RegionT<double>> regions; regions.push_back(make_region<double>(F->fn,S,D,rmin,rmax));
    integrate_regions_sequential<double>(regions,S,D,dimbins,res,rmin,rmax,bins);
    return 0;
}

extern "C" int vo_adaptive_iterations_f64(const char* integrand, const char* rule, const char* heuristic, double size_weight,
                               uint64_t iterations, int dimbins, const uint64_t* res,
                               const double* rmin, const double* rmax, double* bins,
                               double* reg_min, double* reg_max, double* reg_err, uint32_t* reg_dim, double* reg_data) {
    auto F = find_finite_d(integrand); if (!F) return -1;
    int D = F->dim; if (dimbins<1 || dimbins>D) return -2;
    int SH, SL;
    if (!std::strcmp(rule,"simpson_trapezoidal")) { SH=3; SL=2; }
    else if (!std::strcmp(rule,"boole_simpson")) { SH=5; SL=3; }
    else return -2;
    Heuristic h; if (!parse_heuristic(heuristic,size_weight,h)) return -2;
    auto regions = generate_adaptive<double>(F->fn,SH,SL,D,rmin,rmax,h,iterations);
    export_regions<double>(regions,D,reg_min,reg_max,reg_err,reg_dim,reg_data);
    if (bins) integrate_regions_sequential<double>(regions,SH,D,dimbins,res,rmin,rmax,bins);
    return 0;
}

// =========================================================================================================
// Control variates: integrator-crespo2021.h:7-22 -> regions-integrator-parallel-variance-reduction.h:32-109
// =========================================================================================================
namespace {
// RegionsIntegratorParallelVarianceReduction::integrate_regions (regions-integrator-parallel-variance-reduction.h:32-109) with
// rr_uniform_region / cv_optimize_weight / region_sampling_uniform over a D-dimensional region table.
// make_residual(seed) returns the per-bin callable f_regdim(x) (:69).
template<typename MakeResidual>
void cv_integrate_regions(const std::vector<RegionT<float>>& regions, int S, int D, int dimbins, const uint64_t* res,
                          const float* rmin, const float* rmax, uint64_t spp, uint64_t seed, MakeResidual&& make_residual, float* bins,
                          uint32_t* rec_nregions, float* rec_approx, uint32_t* rec_chosen, float* rec_samples,
                          bool fixed_weight = false, double fixed_alpha = 1.0, int rr_policy = 0, int SL = 2) {
    // rr_policy: 0 rr_uniform_region, 1 rr_integral_region, 2 rr_error_region, 3 rr_pdf_region (region-russian-roulette.h:9-28 / :30-67 / :69-106 / :108-147)
    uint64_t nb = nbins_of(dimbins,res);
    uint64_t factor = nb;
    // bin -> region lists, regions visited in list order (:53-57, serial PSTL backend)
    std::vector<std::vector<uint32_t>> perbin(nb);
    for (uint32_t ri=0; ri<regions.size(); ++ri) {
        uint64_t st[8], en[8]; pixels_in_region(regions[ri],dimbins,res,rmin,rmax,st,en);
        uint64_t pos[8]; for (int i=0;i<dimbins;++i) pos[i]=st[i];
        while (true) {
            uint64_t lin=0, prod=1; for (int i=0;i<dimbins;++i) { lin+=pos[i]*prod; prod*=res[i]; }
            perbin[lin].push_back(ri);
            int d=0; for (; d<dimbins; ++d) { if (++pos[d] >= en[d]) pos[d]=st[d]; else break; }
            if (d==dimbins) break;
        }
    }
    // one RNG per bin, seeded sequentially in tensor order from the integrator's mt19937(seed) (:59-63)
    MT19937 master(seed);
    std::vector<uint32_t> bin_seed(nb);
    for (uint64_t k=0;k<nb;++k) bin_seed[k] = master();

    for (uint64_t k=0;k<nb;++k) {                                              // variance_reduction(pos) :67-103
        uint64_t pos[8]; unflatten(k,dimbins,res,pos);
        MT19937 rng{uint64_t(bin_seed[k])};
        // monte_carlo_per_bin(rng,1) seeds itself from the bin RNG: one draw (:69, monte-carlo-per-bin.h:28); what it then does
        // with that seed is the caller's business (nothing for a full-dimensional region table, the rest estimator for Fubini)
        auto f_regdim = make_residual(uint64_t(uint32_t(rng())));
        float ba[8], bb[8]; bin_box(D,dimbins,rmin,rmax,res,pos,ba,bb);       // :71-73
        const auto& list = perbin[k];
        std::size_t n = list.size();
        std::vector<std::array<float,8>> ia(n), ib(n);
        float approximation = 0.0f;                                            // :77
        for (std::size_t i=0;i<n;++i) {                                        // :78-88
            const RegionT<float>& r = regions[list[i]];
            if (intersect(D,ba,bb,r.rmin.data(),r.rmax.data(),ia[i].data(),ib[i].data()))
                approximation = float(double(approximation) + double(factor)*double(region_integral_subrange(r,S,D,ia[i].data(),ib[i].data())));
        }
        if (rec_nregions) rec_nregions[k] = uint32_t(n);
        if (rec_approx) rec_approx[k] = approximation;
        // weighted roulettes: RR constructor (region-russian-roulette.h:42-53 / :81-92) + std::discrete_distribution
        // (libstdc++ bits/random.tcc:2657-2678 _M_initialize: fewer than two weights -> always 0 with probability 1, no draw)
        std::vector<double> prob, cp;
        if (rr_policy != 0) {
            std::vector<double> wts(n);
            for (std::size_t i=0;i<n;++i) {
                const RegionT<float>& r = regions[list[i]];
                if (rr_policy == 1) wts[i] = std::abs(region_integral_subrange(r,S,D,ia[i].data(),ib[i].data()));           // :45 NormDefault = abs
                else if (rr_policy == 2) wts[i] = std::abs(region_error_total(r,S,SL,D))*volume_of(D,ia[i].data(),ib[i].data())/r.volume;   // :86, float arithmetic
                else wts[i] = region_pdf_integral_subrange(r,D,ia[i].data(),ib[i].data());                                     // :125 (Simpson only)
            }
            double sum = 0.0; for (double w : wts) sum += w;
            const float factor_prob = 0.01f;                                    // rr_pdf_region keeps its floor as a FLOAT member (:111)
            if (sum<=0.0) for (double& w : wts) w = 1.0;
            else if (rr_policy == 3) for (double& w : wts) w = std::max(w,factor_prob*sum/double(n));
            else for (double& w : wts) w = std::max(w,0.01*sum/double(n));
            if (n>=2) {
                double total = 0.0; for (double w : wts) total += w;            // std::accumulate(.., 0.0)
                prob.resize(n); cp.resize(n);
                for (std::size_t i=0;i<n;++i) prob[i] = wts[i]/total;           // __normalize
                double run = 0.0; for (std::size_t i=0;i<n;++i) { run = (i==0) ? prob[0] : run + prob[i]; cp[i] = run; }   // std::partial_sum
                cp[n-1] = 1.0;
            }
        }
        // cv_optimize_weight accumulator (weight-strategy.h:40-101)
        float sum_f=0.0f, sum_app=0.0f; uint64_t size=0;
        float fixed_sum = 0.0f;                                                // cv_fixed_weight::Accumulator::sum (weight-strategy.h:14-17)
        double k_f=0,k_app=0,e_f=0,e_ap=0,e_ap2=0,e_fap=0;
        for (uint64_t s=0;s<spp;++s) {                                         // :92-101
            uint64_t chosen; double rrfactor;
            if (rr_policy == 0) { chosen = uniform_int(rng,0,n-1); rrfactor = double(n); }     // region-russian-roulette.h:14,18-21
            else if (cp.empty()) { chosen = 0; rrfactor = 1.0/1.0; }                             // probabilities() == {1.0}
            else {
                double pr = canonical_double(rng);                                               // _Adaptor<URNG,double>
                chosen = uint64_t(std::lower_bound(cp.begin(),cp.end(),pr) - cp.begin());        // bits/random.tcc:2709-2713
                rrfactor = 1.0/prob[chosen];                                                     // region-russian-roulette.h:59
            }
            const RegionT<float>& r = regions[list[chosen]];
            float x[8];
            for (int i=0;i<D;++i) x[i] = uniform_real(rng,ia[chosen][i],ib[chosen][i]);   // region-sampling.h:13-17
            float sfactor = volume_of(D,ia[chosen].data(),ib[chosen].data());              // :18
            if (rec_chosen) rec_chosen[k*spp+s] = list[chosen];
            if (rec_samples) for (int i=0;i<D;++i) rec_samples[(k*spp+s)*D+i] = x[i];
            float fs = float(double(f_regdim(x))*double(factor)*rrfactor*double(sfactor));                      // :98
            float as = float(double(region_approximation_at(r,S,D,x))*double(factor)*rrfactor*double(sfactor)); // :99
            // push (weight-strategy.h:56-69); NormDefault = abs (norm.h:12)
            if (size==0) { k_f = std::abs(fs); k_app = std::abs(as); }
            e_f   += (std::abs(fs) - k_f);
            e_ap  += (std::abs(as) - k_app);
            e_ap2 += (std::abs(as) - k_app)*(std::abs(as) - k_app);
            e_fap += (std::abs(fs) - k_f)*(std::abs(as) - k_app);
            sum_f += fs; sum_app += as;
            fixed_sum = float(double(fixed_sum) + (double(fs) - fixed_alpha*double(as)));      // sum += f - alpha*app  (weight-strategy.h:20)
            ++size;
        }
        // integral (weight-strategy.h:71-101)
        float result;
        if (fixed_weight) result = size==0 ? approximation : float(double(fixed_sum)/double(size) + fixed_alpha*double(approximation));   // :24-27
        else if (size<2) result = approximation;
        else {
            double covariance = (e_fap - (e_f*e_ap)/double(size))/double(size-1);
            double variance   = (e_ap2 - (e_ap*e_ap)/double(size))/double(size-1);
            double alpha;
            double v = std::max(0.0,variance);
            if (v<=0.0) alpha = 1.0;
            else { double c = std::min(v,std::max(0.0,covariance)); alpha = c/v; }
            result = float((double(sum_f) - alpha*double(sum_app))/double(size) + alpha*double(approximation));
        }
        bins[k] = result;                                                      // :102 ('=')
    }
}
} // namespace

extern "C" int vo_crespo2021(const char* integrand, uint64_t iterations, uint64_t spp, uint64_t seed,
                  int dimbins, const uint64_t* res, const float* rmin, const float* rmax, float* bins,
                  uint32_t* rec_nregions, float* rec_approx, uint32_t* rec_chosen, float* rec_samples,
                  float* reg_min, float* reg_max, float* reg_err, uint32_t* reg_dim, float* reg_data) {
    auto F = find_finite(integrand); if (!F) return -1;
    int D = F->dim; if (dimbins<1 || dimbins>D) return -2;
    const int S = 3, SL = 2;
    Heuristic h{true,true,1.e-5};                                              // integrator-crespo2021.h:12
    auto regions = generate_adaptive<float>(F->fn,S,SL,D,rmin,rmax,h,iterations);
    export_regions<float>(regions,D,reg_min,reg_max,reg_err,reg_dim,reg_data);

    cv_integrate_regions(regions,S,D,dimbins,res,rmin,rmax,spp,seed,[&] (uint64_t) { return F->fn; },bins,rec_nregions,rec_approx,rec_chosen,rec_samples);
    return 0;
}

extern "C" int vo_cv_fixed_weight(const char* integrand, uint64_t iterations, uint64_t spp, uint64_t seed, double alpha,
                       int dimbins, const uint64_t* res, const float* rmin, const float* rmax, float* bins,
                       uint32_t* rec_nregions, float* rec_approx, uint32_t* rec_chosen, float* rec_samples) {
    auto F = find_finite(integrand); if (!F) return -1;
    int D = F->dim; if (dimbins<1 || dimbins>D) return -2;
    const int S = 3, SL = 2;
    Heuristic h{true,true,1.e-5};
    auto regions = generate_adaptive<float>(F->fn,S,SL,D,rmin,rmax,h,iterations);
    cv_integrate_regions(regions,S,D,dimbins,res,rmin,rmax,spp,seed,[&] (uint64_t) { return F->fn; },bins,rec_nregions,rec_approx,rec_chosen,rec_samples,true,alpha);
    return 0;
}

extern "C" int vo_cv_policies(const char* integrand, uint64_t iterations, uint64_t spp, uint64_t seed, int rr_policy, int weight_strategy, double alpha,
                   int dimbins, const uint64_t* res, const float* rmin, const float* rmax, float* bins,
                   uint32_t* rec_nregions, float* rec_approx, uint32_t* rec_chosen, float* rec_samples) {
    auto F = find_finite(integrand); if (!F) return -1;
    int D = F->dim; if (dimbins<1 || dimbins>D) return -2;
    if (rr_policy<0 || rr_policy>3 || weight_strategy<0 || weight_strategy>1) return -3;
    const int S = 3, SL = 2;
    Heuristic h{true,true,1.e-5};
    auto regions = generate_adaptive<float>(F->fn,S,SL,D,rmin,rmax,h,iterations);
    cv_integrate_regions(regions,S,D,dimbins,res,rmin,rmax,spp,seed,[&] (uint64_t) { return F->fn; },bins,rec_nregions,rec_approx,rec_chosen,rec_samples,
                         weight_strategy==1,alpha,rr_policy,SL);
    return 0;
}

// =========================================================================================================
// Fubini family: fubini.h:51-101, regions-generator-fubini.h:7-28, integrator-crespo2021.h:24-44
// =========================================================================================================
namespace {

// Every copy of a MonteCarlo / MonteCarloPerBin integrator RESEEDS: new.rng = mt19937(size_t(old.rng()))
// (monte-carlo.h:30-33, monte-carlo-per-bin.h:26-33).  A chain of k copies therefore leaves the last copy seeded by
// s_k, s_{i+1} = first output of mt19937(s_i).  The depths below count the copies between the user's monte_carlo(m, seed)
// and the object that finally evaluates (constructor argument, [=] capture in function_split_and_integrate_at,
// detail::adapt's by-value return, ...): 3 for the rest integrator of integrator_fubini and of regions_generator_fubini, 1 for the
// first integrator of integrator_fubini and for the residual's monte_carlo_per_bin.  They are pinned by the bit-exact tests against
// the unmodified reference (tests/test_oracle_vs_reference.py).
inline uint64_t reseed_chain(uint64_t seed, int depth) {
    for (int i=0;i<depth;++i) { MT19937 g(seed); seed = uint64_t(g()); }
    return seed;
}

// the integrand split at N: finite (f over dim D > N) or a sequence integrand
struct SplitIntegrand {
    const FiniteIntegrand* fin = nullptr; const InfIntegrand* inf = nullptr;
    int N = 0, D = 0;                                   // D: total finite dimension (finite case)
    const float* rmin = nullptr; const float* rmax = nullptr; int nrange = 0;      // full range
    float rest_min(int j) const { int i = N+j; return inf ? (i<nrange ? rmin[i] : 0.0f) : rmin[i]; }      // range_split_at (fubini.h:18-49)
    float rest_max(int j) const { int i = N+j; return inf ? (i<nrange ? rmax[i] : 1.0f) : rmax[i]; }
    int rest_explicit() const { return inf ? std::max(0, nrange-N) : D-N; }
    float rest_volume() const { float v = 1.0f; for (int j=0;j<rest_explicit();++j) v *= (rest_max(j)-rest_min(j)); return v; }
    // f(x ⊕ rest) with the rest elements produced by next(j)
    template<typename Next> float eval(const float* x, Next&& next) const {
        if (!inf) { float full[8]; for (int i=0;i<N;++i) full[i]=x[i]; for (int j=0;j<D-N;++j) full[N+j] = next(j); return fin->fn(full); }
        int idx = 0;
        LazySeq seq;
        seq.next = [&] () -> float { int i = idx++; return i<N ? x[i] : next(i-N); };          // concat(x, xr)  (concat.h:9-45)
        return inf->fn(seq);
    }
};

// g(x) = viltrum::integrate(monte_carlo(m) copy holding `rng`, f(x ⊕ ·), range_rest)   (fubini.h:57-75; monte-carlo.h:39-84, one bin)
inline float rest_monte_carlo(const SplitIntegrand& f, const float* x, uint64_t m, MT19937& rng) {
    double factor = 1.0*double(f.rest_volume())/double(m);                      // monte-carlo.h:43-45 / :70-72
    float sol = 0.0f;
    for (uint64_t s=0;s<m;++s) {
        float v;
        if (!f.inf) {
            float xr[8]; for (int j=0;j<f.D-f.N;++j) xr[j] = uniform_real(rng,f.rest_min(j),f.rest_max(j));     // :50-53
            v = f.eval(x,[&] (int j) { return xr[j]; });
        } else {
            MT19937 seqrng{uint64_t(uint32_t(rng()))};                          // :75 random_sequence(range, seed) -> mt19937(seed)
            v = f.eval(x,[&] (int j) { return uniform_real(seqrng,f.rest_min(j),f.rest_max(j)); });    // random-sequence-rng.h:30,35
        }
        sol = float(double(sol) + double(v)*factor);                            // :59 / :81
    }
    return sol;
}

// g(x) through monte_carlo_per_bin(rng,1) over the rest (regions-integrator-parallel-variance-reduction.h:69; monte-carlo-per-bin.h:41-97, one bin)
inline float rest_monte_carlo_per_bin(const SplitIntegrand& f, const float* x, uint64_t m, MT19937& rng) {
    double factor = double(f.rest_volume())/double(m);
    // the single bin's sub-range rebuilds dimension 0 of the rest: [min + 0*drange, min + 1*drange]
    float a0 = f.rest_min(0), b0 = f.rest_max(0);
    { float drange = (b0-a0)/float(1); float lo = a0 + float(0)*drange, hi = a0 + float(1)*drange; a0 = lo; b0 = hi; }
    auto lo_at = [&] (int j) { return j==0 ? a0 : f.rest_min(j); };
    auto hi_at = [&] (int j) { return j==0 ? b0 : f.rest_max(j); };
    float sol = 0.0f;
    for (uint64_t s=0;s<m;++s) {
        float v;
        if (!f.inf) {
            float xr[8]; for (int j=0;j<f.D-f.N;++j) xr[j] = uniform_real(rng,lo_at(j),hi_at(j));
            v = f.eval(x,[&] (int j) { return xr[j]; });
        } else {
            // Concat::begin() builds the rest iterator at once (concat.h:40), and RandomSequenceRefDis draws its element 0 on
            // construction from the SHARED generator (random-sequence-ref-dis.h:27-28): one draw per sample even when f stops
            // inside the first N elements
            const float e0 = uniform_real(rng,0.0f,1.0f)*(hi_at(0)-lo_at(0))+lo_at(0);
            v = f.eval(x,[&] (int j) { return j==0 ? e0 : uniform_real(rng,0.0f,1.0f)*(hi_at(j)-lo_at(j))+lo_at(j); });   // :28,32
        }
        sol = float(double(sol) + double(v)*factor);
    }
    return sol;
}

int setup_split(const char* integrand, int nfirst, int dimbins, const float* rmin, const float* rmax, int nrange, SplitIntegrand& f) {
    f.fin = find_finite(integrand); f.inf = f.fin ? nullptr : find_inf(integrand);
    if (!f.fin && !f.inf) return -1;
    f.N = nfirst; f.D = f.fin ? f.fin->dim : 0; f.rmin = rmin; f.rmax = rmax; f.nrange = f.fin ? f.D : nrange;
    if (nfirst<1 || nfirst>4 || (f.fin && nfirst>=f.D) || dimbins<1 || dimbins>nfirst || dimbins>2) return -2;
    return 0;
}
// first-part range: the first N entries (implicit [0,1] for infinite ranges)
void first_range(const SplitIntegrand& f, float* a, float* b) {
    for (int i=0;i<f.N;++i) { a[i] = i<f.nrange ? f.rmin[i] : 0.0f; b[i] = i<f.nrange ? f.rmax[i] : 1.0f; }
}

} // namespace

extern "C" int vo_fubini_adaptive_mc(const char* integrand, int nfirst, const char* rule, const char* heuristic, double size_weight,
                          uint64_t iterations, uint64_t mc_samples, uint64_t mc_seed, int dimbins, const uint64_t* res,
                          const float* rmin, const float* rmax, int nrange, float* bins) {
    SplitIntegrand f; int rc = setup_split(integrand,nfirst,dimbins,rmin,rmax,nrange,f); if (rc) return rc;
    int SH, SL;
    if (!std::strcmp(rule,"simpson_trapezoidal")) { SH=3; SL=2; }
    else if (!std::strcmp(rule,"boole_simpson")) { SH=5; SL=3; }
    else return -2;
    Heuristic h; if (!parse_heuristic(heuristic,size_weight,h)) return -2;
    float a[8], b[8]; first_range(f,a,b);
    MT19937 rng(reseed_chain(mc_seed, 3));
    auto g = [&] (const float* x) -> float { return rest_monte_carlo(f,x,mc_samples,rng); };
    auto regions = generate_adaptive<float>(g,SH,SL,f.N,a,b,h,iterations);
    integrate_regions_sequential<float>(regions,SH,f.N,dimbins,res,a,b,bins);
    return 0;
}

extern "C" int vo_fubini_mc_mc(const char* integrand, int nfirst, uint64_t spp, uint64_t seed, uint64_t mc_samples, uint64_t mc_seed,
                    int dimbins, const uint64_t* res, const float* rmin, const float* rmax, int nrange, float* bins) {
    SplitIntegrand f; int rc = setup_split(integrand,nfirst,dimbins,rmin,rmax,nrange,f); if (rc) return rc;
    float ra[8], rb[8]; first_range(f,ra,rb);
    MT19937 rest_rng(reseed_chain(mc_seed, 3));
    // monte_carlo_per_bin_parallel over the first N dimensions (monte-carlo-per-bin-parallel.h:41-71) with f = g
    uint64_t nb = nbins_of(dimbins,res);
    double factor = volume_of(f.N,ra,rb)/double(spp);
    MT19937 master(reseed_chain(seed, 1));          // IntegratorFubini copies the first integrator too (fubini.h:83-84)
    std::vector<uint32_t> perbin_seed(nb);
    for (uint64_t k=0;k<nb;++k) perbin_seed[k] = uint32_t(master());
    // bins are visited in the reference's for_each(parallel) order (serial PSTL backend): dim 0 outer, the rest inside (foreach.h:56-67);
    // the order matters here because the rest integrator's stream runs through all evaluations
    std::vector<uint64_t> order;
    if (dimbins == 1) for (uint64_t k=0;k<nb;++k) order.push_back(k);
    else for (uint64_t d0=0; d0<res[0]; ++d0) for (uint64_t d1=0; d1<res[1]; ++d1) order.push_back(d0 + d1*res[0]);
    for (uint64_t k : order) {
        uint64_t pos[8]; unflatten(k,dimbins,res,pos);
        MT19937 local(perbin_seed[k]);
        float a[8], b[8]; bin_box(f.N,dimbins,ra,rb,res,pos,a,b);
        for (uint64_t s=0;s<spp;++s) {
            float x[8];
            for (int i=0;i<f.N;++i) x[i] = uniform_real(local,a[i],b[i]);
            float v = rest_monte_carlo(f,x,mc_samples,rest_rng);
            bins[k] = float(double(bins[k]) + double(v)*factor);
        }
    }
    return 0;
}

extern "C" int vo_crespo2021_infinite(const char* integrand, int nfirst, uint64_t iterations, uint64_t mc_samples, uint64_t spp, uint64_t seed,
                           int dimbins, const uint64_t* res, const float* rmin, const float* rmax, int nrange, float* bins) {
    SplitIntegrand f; int rc = setup_split(integrand,nfirst,dimbins,rmin,rmax,nrange,f); if (rc) return rc;
    const int S = 3, SL = 2;
    Heuristic h{true,true,1.e-5};                                              // integrator-crespo2021.h:31
    float a[8], b[8]; first_range(f,a,b);
    // generator: monte_carlo(mc_samples, 2*seed+1) (integrator-crespo2021.h:34) copied into RegionsGeneratorFubini, into
    // IntegratorRegionBased and into the [=] closure of function_split_and_integrate_at
    MT19937 gen_rng(reseed_chain(2*seed+1, 3));
    auto g = [&] (const float* x) -> float { return rest_monte_carlo(f,x,mc_samples,gen_rng); };
    auto regions = generate_adaptive<float>(g,S,SL,f.N,a,b,h,iterations);
    // residual: monte_carlo_per_bin(bin rng, 1) captured by value in the closure (one more reseeding copy)
    const int depth = 1;
    cv_integrate_regions(regions,S,f.N,dimbins,res,a,b,spp,seed,[&] (uint64_t s1) {
        auto rng = std::make_shared<MT19937>(reseed_chain(s1, depth));
        return [&f,rng] (const float* x) -> float { return rest_monte_carlo_per_bin(f,x,1,*rng); };
    },bins,nullptr,nullptr,nullptr,nullptr);
    return 0;
}

// Thread-pool driver used as the "port" CPU baseline (same slabbing as the reference-side harness).
extern "C" int vo_mt_per_bin(const char* path, const char* integrand, int dimbins, const uint64_t* res,
                  const float* rmin, const float* rmax, int nrange, uint64_t spp, uint64_t seed,
                  int nthreads, float* bins) {
    if (dimbins < 1 || dimbins > 2) return -2;
    int last = dimbins-1;
    uint64_t rows = res[last];
    if (nthreads < 1) nthreads = 1;
    if (uint64_t(nthreads) > rows) nthreads = int(rows);
    uint64_t stride = (dimbins==2) ? res[0] : 1;
    int dim = vo_integrand_dim(integrand);
    if (dim == 0) return -1;
    bool inf = (dim == -1);
    int have = inf ? nrange : dim;
    int nr = inf ? std::max(nrange, dimbins) : dim;
    std::vector<int> rc(nthreads, 0);
    std::vector<std::thread> th;
    for (int t=0;t<nthreads;++t) th.emplace_back([&,t] () {
        uint64_t lo = rows*uint64_t(t)/uint64_t(nthreads), hi = rows*uint64_t(t+1)/uint64_t(nthreads);
        std::vector<float> a(nr), b(nr);
        for (int i=0;i<nr;++i) { a[i] = (i<have) ? rmin[i] : 0.0f; b[i] = (i<have) ? rmax[i] : 1.0f; }
        float d = (b[last]-a[last])/float(rows);
        float amin = a[last];
        a[last] = amin + float(lo)*d; b[last] = amin + float(hi)*d;
        uint64_t r2[2] = { res[0], res[dimbins>1?1:0] }; r2[last] = hi-lo;
        float* out = bins + lo*stride;
        if (!std::strcmp(path,"mc_per_bin_parallel"))
            rc[t] = vo_mc_per_bin_parallel(integrand,dimbins,r2,a.data(),b.data(),spp,seed+uint64_t(t),out,nullptr,nullptr,nullptr);
        else if (!std::strcmp(path,"per_bin_parallel_mc"))
            rc[t] = vo_per_bin_parallel_mc(integrand,dimbins,r2,a.data(),b.data(),spp,seed+uint64_t(t),out,nullptr,nullptr,nullptr);
        else if (!std::strcmp(path,"mc_per_bin_parallel_inf"))
            rc[t] = vo_mc_per_bin_parallel_inf(integrand,dimbins,r2,a.data(),b.data(),nr,spp,seed+uint64_t(t),out,nullptr,nullptr,nullptr,nullptr,0,nullptr);
        else rc[t] = -2;
        // bins hold densities scaled by the bin count of the call: rescale slab -> full grid
        float scale = float(rows)/float(hi-lo);
        for (uint64_t k=0;k<(hi-lo)*stride;++k) out[k] *= scale;
    });
    for (auto& x : th) x.join();
    for (int x : rc) if (x) return x;
    return 0;
}
