// TEST INFRASTRUCTURE — not product code.
// Newton-Cotes and adaptive (greedy heap) entry points of oracle/_ref/libviltrum_ref.so.
#include "ref_regions.h"

using namespace vref;

namespace { struct MixedArgs { int dimension = 2; double bins_weight = 1.0, tb = 1.0/1024.0, tr = 1.0/16.0, factor = 1.e4; } g_mixed; }
extern "C" void vo_set_mixed(int dimension, double bins_weight, double size_threshold_bins, double size_threshold_rest, double error_increase_factor) {
    g_mixed.dimension = dimension; g_mixed.bins_weight = bins_weight; g_mixed.tb = size_threshold_bins; g_mixed.tr = size_threshold_rest; g_mixed.factor = error_increase_factor;
}

extern "C" int vo_newton_cotes(const char* integrand, const char* rule, int dimbins, const uint64_t* res,
                    const float* rmin, const float* rmax, float* bins) {
    return dispatch_finite_bins(integrand, dimbins, [&] (auto f, auto dbc) -> int {
        using F = decltype(f);
        constexpr std::size_t D = F::dim;
        constexpr std::size_t DB = decltype(dbc)::value;
        auto r = res_array<DB>(res);
        auto range = range_array<D>(rmin, rmax);
        auto acc = [&] (const std::array<std::size_t,DB>& p) -> float& { return bins[tensor_pos(p,r)]; };
        if (!std::strcmp(rule,"trapezoidal")) viltrum::integrate(viltrum::integrator_newton_cotes(viltrum::trapezoidal), acc, r, f, range);
        else if (!std::strcmp(rule,"simpson")) viltrum::integrate(viltrum::integrator_newton_cotes(viltrum::simpson), acc, r, f, range);
        else if (!std::strcmp(rule,"boole"))   viltrum::integrate(viltrum::integrator_newton_cotes(viltrum::boole), acc, r, f, range);
        // composite rules steps<N>(rule) (rules.h:321-388): a fixed menu of instantiations (N is a template argument upstream)
        else if (!std::strcmp(rule,"steps2_boole"))        viltrum::integrate(viltrum::integrator_newton_cotes(viltrum::steps<2>(viltrum::boole)), acc, r, f, range);
        else if (!std::strcmp(rule,"steps3_simpson"))      viltrum::integrate(viltrum::integrator_newton_cotes(viltrum::steps<3>(viltrum::simpson)), acc, r, f, range);
        else if (!std::strcmp(rule,"steps4_trapezoidal"))  viltrum::integrate(viltrum::integrator_newton_cotes(viltrum::steps<4>(viltrum::trapezoidal)), acc, r, f, range);
        else if (!std::strcmp(rule,"steps8_simpson"))      { if constexpr (D <= 3) viltrum::integrate(viltrum::integrator_newton_cotes(viltrum::steps<8>(viltrum::simpson)), acc, r, f, range); else return -2; }
        else if (!std::strcmp(rule,"steps16_trapezoidal")) { if constexpr (D <= 3) viltrum::integrate(viltrum::integrator_newton_cotes(viltrum::steps<16>(viltrum::trapezoidal)), acc, r, f, range); else return -2; }
        else if (!std::strcmp(rule,"steps1_boole"))        viltrum::integrate(viltrum::integrator_newton_cotes(viltrum::steps<1>(viltrum::boole)), acc, r, f, range);
        else return -2;
        return 0;
    });
}

namespace {
template<typename F, std::size_t DB, typename Rule, typename EH>
int run_adaptive(const F& f, const Rule& rule, const EH& eh, uint64_t iterations, const uint64_t* res,
                 const float* rmin, const float* rmax, float* bins, RegionSink& sink) {
    constexpr std::size_t D = F::dim;
    auto r = res_array<DB>(res);
    auto range = range_array<D>(rmin, rmax);
    auto acc = [&] (const std::array<std::size_t,DB>& p) -> float& { return bins[tensor_pos(p,r)]; };
    DumpLogger logger(&sink);
    viltrum::integrate(viltrum::integrator_adaptive_iterations(rule, eh, std::size_t(iterations)), acc, r, f, range, logger);
    return 0;
}
}

extern "C" int vo_adaptive_iterations(const char* integrand, const char* rule, const char* heuristic, double size_weight,
                           uint64_t iterations, int dimbins, const uint64_t* res,
                           const float* rmin, const float* rmax, float* bins,
                           float* reg_min, float* reg_max, float* reg_err, uint32_t* reg_dim, float* reg_data) {
    PhaseTimer phase;
    RegionSink sink; sink.reg_min=reg_min; sink.reg_max=reg_max; sink.reg_err=reg_err; sink.reg_dim=reg_dim; sink.reg_data=reg_data;
    return dispatch_finite_bins(integrand, dimbins, [&] (auto f, auto dbc) -> int {
        constexpr std::size_t DB = decltype(dbc)::value;
        using namespace viltrum;
        auto with_rule = [&] (auto rule) -> int {
            if (!std::strcmp(heuristic,"default_absolute"))
                return run_adaptive<decltype(f),DB>(f, rule, error_heuristic_default(error_metric_absolute()), iterations,res,rmin,rmax,bins,sink);
            if (!std::strcmp(heuristic,"default_relative"))
                return run_adaptive<decltype(f),DB>(f, rule, error_heuristic_default(error_metric_relative()), iterations,res,rmin,rmax,bins,sink);
            if (!std::strcmp(heuristic,"size_absolute"))
                return run_adaptive<decltype(f),DB>(f, rule, error_heuristic_size(error_metric_absolute(),size_weight), iterations,res,rmin,rmax,bins,sink);
            if (!std::strcmp(heuristic,"size_relative"))
                return run_adaptive<decltype(f),DB>(f, rule, error_heuristic_size(error_metric_relative(),size_weight), iterations,res,rmin,rmax,bins,sink);
#define RUN(EH) run_adaptive<decltype(f),DB>(f, rule, EH, iterations,res,rmin,rmax,bins,sink)
            if (!std::strncmp(heuristic,"mixed_",6)) {      // error_heuristic_mixed(bins metric, rest metric, ...) — error-heuristic.h:49-98; extra arguments from vo_set_mixed
                const unsigned dm = unsigned(g_mixed.dimension);
                if (!std::strcmp(heuristic,"mixed_absolute_absolute")) return RUN(error_heuristic_mixed(error_metric_absolute(), error_metric_absolute(), dm, g_mixed.bins_weight, size_weight, g_mixed.tb, g_mixed.tr, g_mixed.factor));
                if (!std::strcmp(heuristic,"mixed_absolute_relative")) return RUN(error_heuristic_mixed(error_metric_absolute(), error_metric_relative(), dm, g_mixed.bins_weight, size_weight, g_mixed.tb, g_mixed.tr, g_mixed.factor));
                if (!std::strcmp(heuristic,"mixed_relative_absolute")) return RUN(error_heuristic_mixed(error_metric_relative(), error_metric_absolute(), dm, g_mixed.bins_weight, size_weight, g_mixed.tb, g_mixed.tr, g_mixed.factor));
                if (!std::strcmp(heuristic,"mixed_relative_relative")) return RUN(error_heuristic_mixed(error_metric_relative(), error_metric_relative(), dm, g_mixed.bins_weight, size_weight, g_mixed.tb, g_mixed.tr, g_mixed.factor));
            }
#undef RUN
            return -2;
        };
        if (!std::strcmp(rule,"simpson_trapezoidal")) return with_rule(nested(simpson,trapezoidal));
        if (!std::strcmp(rule,"boole_simpson"))       return with_rule(nested(boole,simpson));
        return -2;
    });
}

// integrator_adaptive_tolerance — the reference never logs its leaves; every integrated leaf reports progress once
// (integrator-adaptive-tolerance.h:23), which is what the counting logger below counts.
namespace {
struct LeafCountLogger {
    uint64_t* n;
    std::string name() const { return ""; }
    void set_name(const std::string&) {}
    template<typename Number> void log_progress(const Number&, const Number& = Number(1)) { ++*n; }
    template<typename Data> void log(const Data&) {}
};
}
extern "C" int vo_adaptive_tolerance(const char* integrand, const char* rule, const char* heuristic, double size_weight, float tolerance,
                          int dimbins, const uint64_t* res, const float* rmin, const float* rmax, float* bins, uint64_t* nleaves,
                          uint64_t, float*, float*, float*, uint32_t*, float*) {
    uint64_t count = 0;
    int rc = dispatch_finite_bins(integrand, dimbins, [&] (auto f, auto dbc) -> int {
        using F = decltype(f);
        constexpr std::size_t D = F::dim;
        constexpr std::size_t DB = decltype(dbc)::value;
        using namespace viltrum;
        auto r = res_array<DB>(res);
        auto range = range_array<D>(rmin, rmax);
        auto acc = [&] (const std::array<std::size_t,DB>& p) -> float& { return bins[tensor_pos(p,r)]; };
        LeafCountLogger logger{&count};
        auto run = [&] (auto rl, auto eh) -> int { viltrum::integrate(integrator_adaptive_tolerance(rl, eh, tolerance), acc, r, f, range, logger); return 0; };
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
    if (nleaves) *nleaves = count;
    return rc;
}

// ---- double precision: the same reference templates instantiated with Range<double,DIM> ------------------------------------
extern "C" int vo_newton_cotes_f64(const char* integrand, const char* rule, int dimbins, const uint64_t* res,
                        const double* rmin, const double* rmax, double* bins) {
    return dispatch_finite_d(integrand, dimbins, [&] (auto f, auto dbc) -> int {
        using F = decltype(f);
        constexpr std::size_t D = F::dim;
        constexpr std::size_t DB = decltype(dbc)::value;
        auto r = res_array<DB>(res);
        std::array<double,D> a, b; for (std::size_t i=0;i<D;++i) { a[i]=rmin[i]; b[i]=rmax[i]; }
        auto range = viltrum::range(a,b);
        auto acc = [&] (const std::array<std::size_t,DB>& p) -> double& { return bins[tensor_pos(p,r)]; };
        if (!std::strcmp(rule,"trapezoidal")) viltrum::integrate(viltrum::integrator_newton_cotes(viltrum::trapezoidal), acc, r, f, range);
        else if (!std::strcmp(rule,"simpson")) viltrum::integrate(viltrum::integrator_newton_cotes(viltrum::simpson), acc, r, f, range);
        else if (!std::strcmp(rule,"boole"))   viltrum::integrate(viltrum::integrator_newton_cotes(viltrum::boole), acc, r, f, range);
        else return -2;
        return 0;
    });
}

extern "C" int vo_adaptive_iterations_f64(const char* integrand, const char* rule, const char* heuristic, double size_weight,
                               uint64_t iterations, int dimbins, const uint64_t* res,
                               const double* rmin, const double* rmax, double* bins,
                               double* reg_min, double* reg_max, double* reg_err, uint32_t* reg_dim, double* reg_data) {
    RegionSinkT<double> sink; sink.reg_min=reg_min; sink.reg_max=reg_max; sink.reg_err=reg_err; sink.reg_dim=reg_dim; sink.reg_data=reg_data;
    return dispatch_finite_d(integrand, dimbins, [&] (auto f, auto dbc) -> int {
        using F = decltype(f);
        constexpr std::size_t D = F::dim;
        constexpr std::size_t DB = decltype(dbc)::value;
        using namespace viltrum;
        auto r = res_array<DB>(res);
        std::array<double,D> a, b; for (std::size_t i=0;i<D;++i) { a[i]=rmin[i]; b[i]=rmax[i]; }
        auto range = viltrum::range(a,b);
        auto acc = [&] (const std::array<std::size_t,DB>& p) -> double& { return bins[tensor_pos(p,r)]; };
        DumpLoggerT<double> logger(&sink);
        auto run = [&] (auto rl, auto eh) -> int {
            viltrum::integrate(integrator_adaptive_iterations(rl, eh, std::size_t(iterations)), acc, r, f, range, logger);
            return 0;
        };
        auto with_rule = [&] (auto rl) -> int {
            if (!std::strcmp(heuristic,"default_absolute")) return run(rl, error_heuristic_default(error_metric_absolute()));
            if (!std::strcmp(heuristic,"default_relative")) return run(rl, error_heuristic_default(error_metric_relative()));
            if (!std::strcmp(heuristic,"size_absolute"))    return run(rl, error_heuristic_size(error_metric_absolute(),size_weight));
            if (!std::strcmp(heuristic,"size_relative"))    return run(rl, error_heuristic_size(error_metric_relative(),size_weight));
#define RUN(EH) run(rl, EH)
            if (!std::strncmp(heuristic,"mixed_",6)) {      // error_heuristic_mixed(bins metric, rest metric, ...) — error-heuristic.h:49-98; extra arguments from vo_set_mixed
                const unsigned dm = unsigned(g_mixed.dimension);
                if (!std::strcmp(heuristic,"mixed_absolute_absolute")) return RUN(error_heuristic_mixed(error_metric_absolute(), error_metric_absolute(), dm, g_mixed.bins_weight, size_weight, g_mixed.tb, g_mixed.tr, g_mixed.factor));
                if (!std::strcmp(heuristic,"mixed_absolute_relative")) return RUN(error_heuristic_mixed(error_metric_absolute(), error_metric_relative(), dm, g_mixed.bins_weight, size_weight, g_mixed.tb, g_mixed.tr, g_mixed.factor));
                if (!std::strcmp(heuristic,"mixed_relative_absolute")) return RUN(error_heuristic_mixed(error_metric_relative(), error_metric_absolute(), dm, g_mixed.bins_weight, size_weight, g_mixed.tb, g_mixed.tr, g_mixed.factor));
                if (!std::strcmp(heuristic,"mixed_relative_relative")) return RUN(error_heuristic_mixed(error_metric_relative(), error_metric_relative(), dm, g_mixed.bins_weight, size_weight, g_mixed.tb, g_mixed.tr, g_mixed.factor));
            }
#undef RUN
            return -2;
        };
        if (!std::strcmp(rule,"simpson_trapezoidal")) return with_rule(nested(simpson,trapezoidal));
        if (!std::strcmp(rule,"boole_simpson"))       return with_rule(nested(boole,simpson));
        return -2;
    });
}
