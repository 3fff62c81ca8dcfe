// TEST INFRASTRUCTURE — not product code.
// Monte-Carlo entry points of oracle/_ref/libviltrum_ref.so: they call the unmodified reference
// (viltrum::integrate with monte_carlo / monte_carlo_per_bin_parallel / integrator_per_bin_parallel)
// and record what the parity tests need through a wrapping integrand.
#include "ref_common.h"

using namespace vref;

#ifdef VREF_MT
extern "C" const char* vo_kind(void) { return "reference-mt"; }
// threads behind the reference's std::for_each(par_unseq, ...) loops (oracle/pstl_threads/execution); n <= 0 keeps the current value
extern "C" int vo_set_threads(int n) { if (n > 0) vref_pstl::threads() = n; return vref_pstl::threads(); }
#else
extern "C" const char* vo_kind(void) { return "reference"; }
extern "C" int vo_set_threads(int) { return 1; }
#endif
extern "C" void vo_phase_times(double* t) { t[0] = g_phase.t_log - g_phase.t_begin; t[1] = g_phase.t_end - g_phase.t_begin; }

extern "C" int vo_integrand_dim(const char* name) {
    int d = dispatch_finite(name, [] (auto f) -> int { return decltype(f)::dim; });
    if (d > 0) return d;
    if (dispatch_infinite(name, [] (auto) -> int { return 1; }) == 1) return -1;
    return 0;
}

namespace {

// kind 0: monte_carlo_per_bin_parallel ; kind 1: integrator_per_bin_parallel(monte_carlo)
template<int KIND>
int per_bin_finite(const char* integrand, int dimbins, const uint64_t* res,
                   const float* rmin, const float* rmax, uint64_t spp, uint64_t seed,
                   float* bins, float* rec_samples, double* rec_sum, double* rec_sum2) {
    return dispatch_finite_bins(integrand, dimbins, [&] (auto f, auto dbc) -> int {
        using F = decltype(f);
        constexpr std::size_t D = F::dim;
        constexpr std::size_t DB = decltype(dbc)::value;
        auto r = res_array<DB>(res);
        auto range = range_array<D>(rmin, rmax);
        auto order = parallel_visit_order<DB>(r);
        std::size_t calls = 0;
        auto recf = [&] (const std::array<float,D>& x) -> float {
            float v = f(x);
            std::size_t k = calls++;
            std::size_t bin = order[k/spp];
            std::size_t s = k % spp;
            if (rec_samples) for (std::size_t i=0;i<D;++i) rec_samples[(bin*spp+s)*D+i] = x[i];
            if (rec_sum)  rec_sum[bin]  += double(v);
            if (rec_sum2) rec_sum2[bin] += double(v)*double(v);
            return v;
        };
        auto acc = [&] (const std::array<std::size_t,DB>& p) -> float& { return bins[tensor_pos(p,r)]; };
        std::size_t nbins = 1; for (auto x : r) nbins *= x;
        if (rec_sum)  std::fill(rec_sum,  rec_sum+nbins,  0.0);
        if (rec_sum2) std::fill(rec_sum2, rec_sum2+nbins, 0.0);
        if (!rec_samples && !rec_sum && !rec_sum2) {      // nothing to record (the timing legs): the integrand itself, no wrapper
            if constexpr (KIND == 0)
                viltrum::integrate(viltrum::monte_carlo_per_bin_parallel(spp, std::size_t(seed)), acc, r, f, range);
            else
                viltrum::integrate(viltrum::integrator_per_bin_parallel(viltrum::monte_carlo(spp, std::size_t(seed))), acc, r, f, range);
            return 0;
        }
#ifdef VREF_MT
        // multi-threaded build: calls arrive in no particular order, so the bin of a sample is found from its coordinates (a sample that
        // float rounding puts exactly on a bin's upper edge is booked to the neighbour: ~1e-7 of the samples, irrelevant for the moments)
        if (rec_samples) return -4;
        {
            float rmn[DB], inv[DB];
            for (std::size_t i=0;i<DB;++i) { rmn[i] = rmin[i]; inv[i] = float(r[i])/(rmax[i]-rmin[i]); }
            auto posf = [&] (const std::array<float,D>& x) -> float {
                float v = f(x);
                std::size_t bin = 0, prod = 1;
                for (std::size_t i=0;i<DB;++i) { std::size_t k = std::size_t((x[i]-rmn[i])*inv[i]); if (k >= r[i]) k = r[i]-1; bin += k*prod; prod *= r[i]; }
                if (rec_sum)  rec_sum[bin]  += double(v);
                if (rec_sum2) rec_sum2[bin] += double(v)*double(v);
                return v;
            };
            if constexpr (KIND == 0)
                viltrum::integrate(viltrum::monte_carlo_per_bin_parallel(spp, std::size_t(seed)), acc, r, posf, range);
            else
                viltrum::integrate(viltrum::integrator_per_bin_parallel(viltrum::monte_carlo(spp, std::size_t(seed))), acc, r, posf, range);
            return 0;
        }
#endif
        if constexpr (KIND == 0)
            viltrum::integrate(viltrum::monte_carlo_per_bin_parallel(spp, std::size_t(seed)), acc, r, recf, range);
        else
            viltrum::integrate(viltrum::integrator_per_bin_parallel(viltrum::monte_carlo(spp, std::size_t(seed))), acc, r, recf, range);
        return 0;
    });
}

} // namespace

extern "C" int vo_mc_per_bin_parallel(const char* integrand, int dimbins, const uint64_t* res,
                           const float* rmin, const float* rmax, uint64_t spp, uint64_t seed,
                           float* bins, float* rec_samples, double* rec_sum, double* rec_sum2) {
    return per_bin_finite<0>(integrand,dimbins,res,rmin,rmax,spp,seed,bins,rec_samples,rec_sum,rec_sum2);
}

extern "C" int vo_per_bin_parallel_mc(const char* integrand, int dimbins, const uint64_t* res,
                           const float* rmin, const float* rmax, uint64_t spp, uint64_t seed,
                           float* bins, float* rec_samples, double* rec_sum, double* rec_sum2) {
    return per_bin_finite<1>(integrand,dimbins,res,rmin,rmax,spp,seed,bins,rec_samples,rec_sum,rec_sum2);
}

extern "C" int vo_monte_carlo(const char* integrand, int dimbins, const uint64_t* res,
                   const float* rmin, const float* rmax, uint64_t samples, uint64_t seed,
                   float* bins, float* rec_samples) {
    return dispatch_finite_bins(integrand, dimbins, [&] (auto f, auto dbc) -> int {
        using F = decltype(f);
        constexpr std::size_t D = F::dim;
        constexpr std::size_t DB = decltype(dbc)::value;
        auto r = res_array<DB>(res);
        auto range = range_array<D>(rmin, rmax);
        std::size_t calls = 0;
        auto recf = [&] (const std::array<float,D>& x) -> float {
            std::size_t k = calls++;
            if (rec_samples) for (std::size_t i=0;i<D;++i) rec_samples[k*D+i] = x[i];
            return f(x);
        };
        auto acc = [&] (const std::array<std::size_t,DB>& p) -> float& { return bins[tensor_pos(p,r)]; };
        viltrum::integrate(viltrum::monte_carlo(samples, std::size_t(seed)), acc, r, recf, range);
        return 0;
    });
}

namespace {
// Sequence wrapper that records every element the underlying lazy sequence generates (begin() already
// draws element 0, each ++ draws the next: reference src/monte-carlo/random-sequence-ref-dis.h:27-32).
struct SeqRecorder { std::vector<float>* elems; uint32_t* count; };
template<typename Seq>
class RecSeq {
    const Seq& seq; SeqRecorder rec;
public:
    RecSeq(const Seq& s, SeqRecorder r) : seq(s), rec(r) {}
    class const_iterator {
        decltype(std::declval<const Seq&>().begin()) it; SeqRecorder rec;
        void note() { if (rec.elems) rec.elems->push_back(*it); if (rec.count) ++(*rec.count); }
    public:
        const_iterator(const Seq& s, SeqRecorder r) : it(s.begin()), rec(r) { note(); }
        const float& operator*() const { return *it; }
        const_iterator& operator++() { ++it; note(); return *this; }
    };
    const_iterator begin() const { return const_iterator(seq, rec); }
};
}

extern "C" int vo_mc_per_bin_parallel_inf(const char* integrand, int dimbins, const uint64_t* res,
                               const float* rmin, const float* rmax, int nrange,
                               uint64_t spp, uint64_t seed, float* bins,
                               double* rec_sum, double* rec_sum2,
                               uint32_t* rec_len, float* rec_elems, uint64_t rec_cap, uint64_t* rec_used) {
    return dispatch_infinite(integrand, [&] (auto f) -> int {
        auto run = [&] (auto dbc) -> int {
            constexpr std::size_t DB = decltype(dbc)::value;
            auto r = res_array<DB>(res);
            auto range = viltrum::range_infinite(std::vector<float>(rmin,rmin+nrange), std::vector<float>(rmax,rmax+nrange));
            auto order = parallel_visit_order<DB>(r);
            std::size_t nbins = 1; for (auto x : r) nbins *= x;
            if (rec_sum)  std::fill(rec_sum,  rec_sum+nbins,  0.0);
            if (rec_sum2) std::fill(rec_sum2, rec_sum2+nbins, 0.0);
            // elements are recorded per path in visit order, then permuted to tensor order at the end
            std::vector<std::vector<float>> per_path(rec_elems || rec_used ? nbins*spp : 0);
            std::vector<uint32_t> lens(nbins*spp, 0);
            std::size_t calls = 0;
            auto recf = [&] (const auto& seq) -> float {
                std::size_t k = calls++;
                std::size_t bin = order[k/spp]; std::size_t s = k % spp;
                SeqRecorder sr{ per_path.empty() ? nullptr : &per_path[bin*spp+s], &lens[bin*spp+s] };
                RecSeq<std::decay_t<decltype(seq)>> rs(seq, sr);
                float v = f(rs);
                if (rec_sum)  rec_sum[bin]  += double(v);
                if (rec_sum2) rec_sum2[bin] += double(v)*double(v);
                return v;
            };
            auto acc = [&] (const std::array<std::size_t,DB>& p) -> float& { return bins[tensor_pos(p,r)]; };
            if (!rec_sum && !rec_sum2 && !rec_len && !rec_elems) {      // timing legs: the integrand itself
                viltrum::integrate(viltrum::monte_carlo_per_bin_parallel(spp, std::size_t(seed)), acc, r, f, range);
                if (rec_used) *rec_used = 0;
                return 0;
            }
#ifdef VREF_MT
            // multi-threaded build: per-bin moments only, the bin found from the first DB elements the integrand itself consumed
            if (rec_len || rec_elems) return -4;
            {
                float rmn[DB], inv[DB];
                for (std::size_t i=0;i<DB;++i) { float a = i<std::size_t(nrange)?rmin[i]:0.0f, b = i<std::size_t(nrange)?rmax[i]:1.0f; rmn[i] = a; inv[i] = float(r[i])/(b-a); }
                auto posf = [&] (const auto& seq) -> float {
                    std::vector<float> seen; seen.reserve(16);
                    RecSeq<std::decay_t<decltype(seq)>> rs(seq, SeqRecorder{&seen, nullptr});
                    float v = f(rs);
                    std::size_t bin = 0, prod = 1;
                    for (std::size_t i=0;i<DB;++i) { std::size_t k = i<seen.size() ? std::size_t((seen[i]-rmn[i])*inv[i]) : 0; if (k >= r[i]) k = r[i]-1; bin += k*prod; prod *= r[i]; }
                    if (rec_sum)  rec_sum[bin]  += double(v);
                    if (rec_sum2) rec_sum2[bin] += double(v)*double(v);
                    return v;
                };
                viltrum::integrate(viltrum::monte_carlo_per_bin_parallel(spp, std::size_t(seed)), acc, r, posf, range);
                if (rec_used) *rec_used = 0;
                return 0;
            }
#endif
            viltrum::integrate(viltrum::monte_carlo_per_bin_parallel(spp, std::size_t(seed)), acc, r, recf, range);
            if (rec_len) std::copy(lens.begin(), lens.end(), rec_len);
            uint64_t used = 0; for (auto l : lens) used += l;
            if (rec_used) *rec_used = used;
            if (rec_elems) {
                if (used > rec_cap) return -3;
                uint64_t o = 0;
                for (auto& v : per_path) { std::copy(v.begin(), v.end(), rec_elems+o); o += v.size(); }
            }
            return 0;
        };
        if (dimbins == 1) return run(std::integral_constant<std::size_t,1>());
        if (dimbins == 2) return run(std::integral_constant<std::size_t,2>());
        return -2;
    });
}

extern "C" int vo_per_bin_parallel_mc_inf(const char* integrand, int dimbins, const uint64_t* res,
                               const float* rmin, const float* rmax, int nrange,
                               uint64_t spp, uint64_t seed, float* bins,
                               double* rec_sum, double* rec_sum2,
                               uint32_t* rec_len, float* rec_elems, uint64_t rec_cap, uint64_t* rec_used) {
    return dispatch_infinite(integrand, [&] (auto f) -> int {
        auto run = [&] (auto dbc) -> int {
            constexpr std::size_t DB = decltype(dbc)::value;
            auto r = res_array<DB>(res);
            auto range = viltrum::range_infinite(std::vector<float>(rmin,rmin+nrange), std::vector<float>(rmax,rmax+nrange));
            auto order = parallel_visit_order<DB>(r);
            std::size_t nbins = 1; for (auto x : r) nbins *= x;
            if (rec_sum)  std::fill(rec_sum,  rec_sum+nbins,  0.0);
            if (rec_sum2) std::fill(rec_sum2, rec_sum2+nbins, 0.0);
            std::vector<std::vector<float>> per_path(rec_elems || rec_used ? nbins*spp : 0);
            std::vector<uint32_t> lens(nbins*spp, 0);
            std::size_t calls = 0;
            // the value-returning integrate deduces T from function(std::vector<Float>()) (integrate.h:117): the recorder must
            // only count REAL calls, which come with the reference's RandomSequenceRNG
            auto recf = [&] (const auto& seq) -> float {
                using Seq = std::decay_t<decltype(seq)>;
                if constexpr (std::is_same_v<Seq, std::vector<float>>) { return f(seq); }
                else {
                    std::size_t k = calls++;
                    std::size_t bin = order[k/spp]; std::size_t s = k % spp;
                    SeqRecorder sr{ per_path.empty() ? nullptr : &per_path[bin*spp+s], &lens[bin*spp+s] };
                    RecSeq<Seq> rs(seq, sr);
                    float v = f(rs);
                    if (rec_sum)  rec_sum[bin]  += double(v);
                    if (rec_sum2) rec_sum2[bin] += double(v)*double(v);
                    return v;
                }
            };
            auto acc = [&] (const std::array<std::size_t,DB>& p) -> float& { return bins[tensor_pos(p,r)]; };
            viltrum::integrate(viltrum::integrator_per_bin_parallel(viltrum::monte_carlo(spp, std::size_t(seed))), acc, r, recf, range);
            if (rec_len) std::copy(lens.begin(), lens.end(), rec_len);
            uint64_t used = 0; for (auto l : lens) used += l;
            if (rec_used) *rec_used = used;
            if (rec_elems) {
                if (used > rec_cap) return -3;
                uint64_t o = 0;
                for (auto& v : per_path) { std::copy(v.begin(), v.end(), rec_elems+o); o += v.size(); }
            }
            return 0;
        };
        if (dimbins == 1) return run(std::integral_constant<std::size_t,1>());
        if (dimbins == 2) return run(std::integral_constant<std::size_t,2>());
        return -2;
    });
}

extern "C" int vo_monte_carlo_inf(const char* integrand, int dimbins, const uint64_t* res,
                       const float* rmin, const float* rmax, int nrange, uint64_t samples, uint64_t seed, float* bins) {
    return dispatch_infinite(integrand, [&] (auto f) -> int {
        auto run = [&] (auto dbc) -> int {
            constexpr std::size_t DB = decltype(dbc)::value;
            auto r = res_array<DB>(res);
            auto range = viltrum::range_infinite(std::vector<float>(rmin,rmin+nrange), std::vector<float>(rmax,rmax+nrange));
            auto acc = [&] (const std::array<std::size_t,DB>& p) -> float& { return bins[tensor_pos(p,r)]; };
            viltrum::integrate(viltrum::monte_carlo(samples, std::size_t(seed)), acc, r, f, range);
            return 0;
        };
        if (dimbins == 1) return run(std::integral_constant<std::size_t,1>());
        if (dimbins == 2) return run(std::integral_constant<std::size_t,2>());
        return -2;
    });
}

// Thread-pool CPU baseline: slab the LAST bin dimension over std::threads; each slab is an
// independent call of the unmodified single-threaded reference with a sub-range and seed+slab.
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
    bool inf = (dim == -1);
    if (dim == 0) return -1;
    int nr = inf ? std::max(nrange, dimbins) : dim;
    std::vector<int> rc(nthreads, 0);
    std::vector<std::thread> th;
    for (int t=0;t<nthreads;++t) th.emplace_back([&,t] () {
        uint64_t lo = rows*uint64_t(t)/uint64_t(nthreads), hi = rows*uint64_t(t+1)/uint64_t(nthreads);
        std::vector<float> a(nr), b(nr);
        for (int i=0;i<nr;++i) { a[i] = (i<(inf?nrange:dim)) ? rmin[i] : 0.0f; b[i] = (i<(inf?nrange:dim)) ? rmax[i] : 1.0f; }
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
