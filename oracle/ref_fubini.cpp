// TEST INFRASTRUCTURE — not product code.
// Fubini-family entry points of oracle/_ref/libviltrum_ref.so: the unmodified reference's integrator_fubini<N>
// (src/combination/fubini.h:78-101) and integrator_crespo2021_infinite<N> (src/control-variates/integrator-crespo2021.h:24-44),
// over finite ranges (rest = the integrand's remaining dimensions) and infinite ones (rest = a lazy sequence).
#include "ref_regions.h"

using namespace vref;

namespace {

template<std::size_t DB>
struct Acc {
    float* bins; std::array<std::size_t,DB> r;
    float& operator()(const std::array<std::size_t,DB>& p) const { return bins[tensor_pos(p,r)]; }
};

// fn(f, range, N-constant, DB-constant) for the named integrand / split / bin dimensionality
template<typename Fn>
int dispatch_split(const char* integrand, int nfirst, int dimbins, const float* rmin, const float* rmax, int nrange, Fn&& fn) {
    auto with_n_db = [&] (auto f, auto range, auto nc) -> int {
        constexpr std::size_t N = decltype(nc)::value;
        if (dimbins == 1) return fn(f, range, nc, std::integral_constant<std::size_t,1>());
        if constexpr (N >= 2) { if (dimbins == 2) return fn(f, range, nc, std::integral_constant<std::size_t,2>()); }
        return -2;
    };
    int rc = dispatch_finite(integrand, [&] (auto f) -> int {
        using F = decltype(f);
        constexpr std::size_t D = F::dim;
        auto range = range_array<D>(rmin, rmax);
        if constexpr (D > 1) { if (nfirst == 1) return with_n_db(f, range, std::integral_constant<std::size_t,1>()); }
        if constexpr (D > 2) { if (nfirst == 2) return with_n_db(f, range, std::integral_constant<std::size_t,2>()); }
        if constexpr (D > 3) { if (nfirst == 3) return with_n_db(f, range, std::integral_constant<std::size_t,3>()); }
        if constexpr (D > 4) { if (nfirst == 4) return with_n_db(f, range, std::integral_constant<std::size_t,4>()); }
        return -2;
    });
    if (rc != -1) return rc;
    return dispatch_infinite(integrand, [&] (auto f) -> int {
        auto range = viltrum::range_infinite(std::vector<float>(rmin, rmin+nrange), std::vector<float>(rmax, rmax+nrange));
        if (nfirst == 1) return with_n_db(f, range, std::integral_constant<std::size_t,1>());
        if (nfirst == 2) return with_n_db(f, range, std::integral_constant<std::size_t,2>());
        if (nfirst == 3) return with_n_db(f, range, std::integral_constant<std::size_t,3>());
        if (nfirst == 4) return with_n_db(f, range, std::integral_constant<std::size_t,4>());
        return -2;
    });
}

} // namespace

extern "C" int vo_fubini_adaptive_mc(const char* integrand, int nfirst, const char* rule, const char* heuristic, double size_weight,
                          uint64_t iterations, uint64_t mc_samples, uint64_t mc_seed, int dimbins, const uint64_t* res,
                          const float* rmin, const float* rmax, int nrange, float* bins) {
    return dispatch_split(integrand, nfirst, dimbins, rmin, rmax, nrange, [&] (auto f, auto range, auto nc, auto dbc) -> int {
        constexpr std::size_t N = decltype(nc)::value;
        constexpr std::size_t DB = decltype(dbc)::value;
        using namespace viltrum;
        Acc<DB> acc{bins, res_array<DB>(res)};
        auto run = [&] (auto rl, auto eh) -> int {
            viltrum::integrate(integrator_fubini<N>(integrator_adaptive_iterations(rl, eh, std::size_t(iterations)),
                                                    monte_carlo((unsigned long)mc_samples, std::size_t(mc_seed))), acc, acc.r, f, range);
            return 0;
        };
        auto with_rule = [&] (auto rl) -> int {
            if (!std::strcmp(heuristic,"default_absolute")) return run(rl, error_heuristic_default(error_metric_absolute()));
            if (!std::strcmp(heuristic,"default_relative")) return run(rl, error_heuristic_default(error_metric_relative()));
            if (!std::strcmp(heuristic,"size_absolute"))    return run(rl, error_heuristic_size(error_metric_absolute(),size_weight));
            if (!std::strcmp(heuristic,"size_relative"))    return run(rl, error_heuristic_size(error_metric_relative(),size_weight));
            return -2;
        };
        if (!std::strcmp(rule,"simpson_trapezoidal")) return with_rule(nested(simpson,trapezoidal));
        if (!std::strcmp(rule,"boole_simpson"))       return with_rule(nested(boole,simpson));
        return -2;
    });
}

extern "C" int vo_fubini_mc_mc(const char* integrand, int nfirst, uint64_t spp, uint64_t seed, uint64_t mc_samples, uint64_t mc_seed,
                    int dimbins, const uint64_t* res, const float* rmin, const float* rmax, int nrange, float* bins) {
    return dispatch_split(integrand, nfirst, dimbins, rmin, rmax, nrange, [&] (auto f, auto range, auto nc, auto dbc) -> int {
        constexpr std::size_t N = decltype(nc)::value;
        constexpr std::size_t DB = decltype(dbc)::value;
        using namespace viltrum;
        Acc<DB> acc{bins, res_array<DB>(res)};
        viltrum::integrate(integrator_fubini<N>(monte_carlo_per_bin_parallel((unsigned long)spp, std::size_t(seed)),
                                                monte_carlo((unsigned long)mc_samples, std::size_t(mc_seed))), acc, acc.r, f, range);
        return 0;
    });
}

extern "C" int vo_crespo2021_infinite(const char* integrand, int nfirst, uint64_t iterations, uint64_t mc_samples, uint64_t spp, uint64_t seed,
                           int dimbins, const uint64_t* res, const float* rmin, const float* rmax, int nrange, float* bins) {
    return dispatch_split(integrand, nfirst, dimbins, rmin, rmax, nrange, [&] (auto f, auto range, auto nc, auto dbc) -> int {
        constexpr std::size_t N = decltype(nc)::value;
        constexpr std::size_t DB = decltype(dbc)::value;
        Acc<DB> acc{bins, res_array<DB>(res)};
        viltrum::integrate(viltrum::integrator_crespo2021_infinite<N>(std::size_t(iterations), std::size_t(mc_samples), std::size_t(spp), std::size_t(seed)),
                           acc, acc.r, f, range);
        return 0;
    });
}


// reference integrator_adaptive_fubini_variance_reduction_parallel_optimized<N>(nested(simpson,trapezoidal), error_heuristic_size(relative,1e-5),
// iterations, mc_samples, region_stratification_uniform(), cv_optimize_weight(), region_sampling_uniform(), spp, seed) —
// src/control-variates/integrator-adaptive-fubini-variance-reduction-optimized.h:17-23 -> RegionsIntegratorParallelVarianceReductionOptimized
// (…-variance-reduction-optimized.h:70-142).  Infinite ranges only: over a finite rest the class does not compile upstream (Range::DIM, :27).  '='.
extern "C" int vo_cv_optimized_infinite(const char* integrand, int nfirst, uint64_t iterations, uint64_t mc_samples, uint64_t spp, uint64_t seed,
                             int dimbins, const uint64_t* res, const float* rmin, const float* rmax, int nrange, float* bins) {
    return dispatch_infinite(integrand, [&] (auto f) -> int {
        auto range = viltrum::range_infinite(std::vector<float>(rmin, rmin+nrange), std::vector<float>(rmax, rmax+nrange));
        auto go = [&] (auto nc, auto dbc) -> int {
            constexpr std::size_t N = decltype(nc)::value;
            constexpr std::size_t DB = decltype(dbc)::value;
            using namespace viltrum;
            Acc<DB> acc{bins, res_array<DB>(res)};
            viltrum::integrate(integrator_adaptive_fubini_variance_reduction_parallel_optimized<N>(nested(simpson,trapezoidal), error_heuristic_size(error_metric_relative(),1.e-5),
                                   std::size_t(iterations), (unsigned long)mc_samples, region_stratification_uniform(), cv_optimize_weight(), region_sampling_uniform(),
                                   (unsigned long)spp, std::size_t(seed)), acc, acc.r, f, range);
            return 0;
        };
        if (nfirst == 1 && dimbins == 1) return go(std::integral_constant<std::size_t,1>(), std::integral_constant<std::size_t,1>());
        if (nfirst == 2 && dimbins == 1) return go(std::integral_constant<std::size_t,2>(), std::integral_constant<std::size_t,1>());
        if (nfirst == 2 && dimbins == 2) return go(std::integral_constant<std::size_t,2>(), std::integral_constant<std::size_t,2>());
        if (nfirst == 3 && dimbins == 2) return go(std::integral_constant<std::size_t,3>(), std::integral_constant<std::size_t,2>());
        return -2;
    });
}
