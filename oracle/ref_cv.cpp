// TEST INFRASTRUCTURE — not product code.
// Control-variate entry point of oracle/_ref/libviltrum_ref.so: the reference's integrator_crespo2021
// preset (src/control-variates/integrator-crespo2021.h:7-22) rebuilt from the same public factories, with
// recording wrappers passed as the RR / CV / RS policy template arguments so the per-sample region choices,
// sample points and the per-bin control-variate integral can be captured without touching the reference.
#include "ref_regions.h"

using namespace vref;

namespace {

struct CvRecorder {
    int db = 0; float rmin[2] = {0,0}; float drange[2] = {1,1}; std::size_t res[2] = {1,1};
    uint64_t spp = 0; std::size_t dim = 0;
    uint32_t* nregions = nullptr; float* approx = nullptr; uint32_t* chosen = nullptr; float* samples = nullptr;
    const RegionSink* sink = nullptr; std::size_t reg_size = 0;
    std::size_t cur_bin = 0, cur_sample = 0;

    template<typename R> void begin_bin(const R& range, std::size_t n) {
        std::size_t pos = 0, prod = 1;
        for (int i=0;i<db;++i) {
            float mid = 0.5f*(range.min(i)+range.max(i));
            std::size_t k = std::size_t((mid - rmin[i])/drange[i]);
            if (k >= res[i]) k = res[i]-1;
            pos += k*prod; prod *= res[i];
        }
        cur_bin = pos; cur_sample = 0;
        if (nregions) nregions[pos] = uint32_t(n);
    }
};

template<typename Inner>
class RecRRT {
    CvRecorder* rec; Inner inner;
public:
    RecRRT(CvRecorder* r, const Inner& i = Inner()) : rec(r), inner(i) {}
    class RR {
        typename Inner::RR inner;
    public:
        RR(typename Inner::RR&& i) : inner(std::move(i)) {}
        template<typename RNG> std::tuple<std::size_t,double> choose(RNG& rng) { return inner.choose(rng); }
    };
    template<typename Regions> RR russian_roulette(const Regions& regions) const {
        if (!regions.empty()) rec->begin_bin(std::get<1>(regions[0]), regions.size());
        return RR(inner.russian_roulette(regions));
    }
};
using RecRR = RecRRT<viltrum::rr_uniform_region>;

class RecRS {
    CvRecorder* rec; viltrum::region_sampling_uniform inner;
public:
    RecRS(CvRecorder* r) : rec(r) {}
    template<typename R, typename Float, std::size_t DIM, typename RNG>
    std::tuple<std::array<Float,DIM>,Float> sample(const R* reg, const viltrum::Range<Float,DIM>& range, RNG& rng) const {
        auto t = inner.sample(reg, range, rng);
        std::size_t k = rec->cur_bin*rec->spp + rec->cur_sample;
        if (rec->chosen) rec->chosen[k] = uint32_t((reinterpret_cast<const char*>(reg) - static_cast<const char*>(rec->sink->base))/sizeof(R));
        if (rec->samples) for (std::size_t i=0;i<DIM;++i) rec->samples[k*DIM+i] = std::get<0>(t)[i];
        ++rec->cur_sample;
        return t;
    }
};

class RecCV {
    CvRecorder* rec; viltrum::cv_optimize_weight<> inner;
public:
    RecCV(CvRecorder* r) : rec(r) {}
    template<typename Sample> class Accumulator {
        typename viltrum::cv_optimize_weight<>::template Accumulator<Sample> acc; CvRecorder* rec;
    public:
        Accumulator(typename viltrum::cv_optimize_weight<>::template Accumulator<Sample>&& a, CvRecorder* r) : acc(std::move(a)), rec(r) {}
        void push(const Sample& f, const Sample& a) { acc.push(f,a); }
        Sample integral(const Sample& approximation) const {
            if (rec->approx) rec->approx[rec->cur_bin] = approximation;
            return acc.integral(approximation);
        }
    };
    template<typename Sample> Accumulator<Sample> accumulator(const Sample& ini = Sample(0)) const {
        return Accumulator<Sample>(inner.template accumulator<Sample>(ini), rec);
    }
};

} // namespace

extern "C" int vo_crespo2021(const char* integrand, uint64_t iterations, uint64_t spp, uint64_t seed,
                  int dimbins, const uint64_t* res, const float* rmin, const float* rmax, float* bins,
                  uint32_t* rec_nregions, float* rec_approx, uint32_t* rec_chosen, float* rec_samples,
                  float* reg_min, float* reg_max, float* reg_err, uint32_t* reg_dim, float* reg_data) {
    PhaseTimer phase;
    RegionSink sink; sink.reg_min=reg_min; sink.reg_max=reg_max; sink.reg_err=reg_err; sink.reg_dim=reg_dim; sink.reg_data=reg_data;
    return dispatch_finite_bins(integrand, dimbins, [&] (auto f, auto dbc) -> int {
        using F = decltype(f);
        constexpr std::size_t D = F::dim;
        constexpr std::size_t DB = decltype(dbc)::value;
        using namespace viltrum;
        auto r = res_array<DB>(res);
        auto range = range_array<D>(rmin, rmax);
        auto acc = [&] (const std::array<std::size_t,DB>& p) -> float& { return bins[tensor_pos(p,r)]; };
        CvRecorder rec; rec.db = int(DB); rec.spp = spp; rec.dim = D; rec.sink = &sink;
        for (std::size_t i=0;i<DB;++i) { rec.rmin[i]=rmin[i]; rec.res[i]=r[i]; rec.drange[i]=(rmax[i]-rmin[i])/float(r[i]); }
        rec.nregions=rec_nregions; rec.approx=rec_approx; rec.chosen=rec_chosen; rec.samples=rec_samples;
        DumpLogger logger(&sink);
        bool record = rec_nregions || rec_approx || rec_chosen || rec_samples;
        if (record) {
            // same factories and arguments as integrator_crespo2021 (integrator-crespo2021.h:9-21), policies wrapped
            auto integrator = integrator_region_based(
                regions_generator_adaptive_heap(nested(simpson,trapezoidal), error_heuristic_size(error_metric_relative(),1.e-5), std::size_t(iterations)),
                regions_integrator_parallel_variance_reduction(RecRR(&rec), RecCV(&rec), RecRS(&rec), std::mt19937(std::size_t(seed)), (unsigned long)spp, std::size_t(16)));
            viltrum::integrate(integrator, acc, r, f, range, logger);
        } else {
            viltrum::integrate(integrator_crespo2021(std::size_t(iterations), std::size_t(spp), std::size_t(seed)), acc, r, f, range, logger);
        }
        return 0;
    });
}

extern "C" int vo_cv_fixed_weight(const char* integrand, uint64_t iterations, uint64_t spp, uint64_t seed, double alpha,
                       int dimbins, const uint64_t* res, const float* rmin, const float* rmax, float* bins,
                       uint32_t* rec_nregions, float* rec_approx, uint32_t* rec_chosen, float* rec_samples) {
    RegionSink sink;
    return dispatch_finite_bins(integrand, dimbins, [&] (auto f, auto dbc) -> int {
        using F = decltype(f);
        constexpr std::size_t D = F::dim;
        constexpr std::size_t DB = decltype(dbc)::value;
        using namespace viltrum;
        auto r = res_array<DB>(res);
        auto range = range_array<D>(rmin, rmax);
        auto acc = [&] (const std::array<std::size_t,DB>& p) -> float& { return bins[tensor_pos(p,r)]; };
        CvRecorder rec; rec.db = int(DB); rec.spp = spp; rec.dim = D; rec.sink = &sink;
        for (std::size_t i=0;i<DB;++i) { rec.rmin[i]=rmin[i]; rec.res[i]=r[i]; rec.drange[i]=(rmax[i]-rmin[i])/float(r[i]); }
        rec.nregions=rec_nregions; rec.approx=nullptr; rec.chosen=rec_chosen; rec.samples=rec_samples;
        DumpLogger logger(&sink);
        auto integrator = integrator_region_based(
            regions_generator_adaptive_heap(nested(simpson,trapezoidal), error_heuristic_size(error_metric_relative(),1.e-5), std::size_t(iterations)),
            regions_integrator_parallel_variance_reduction(RecRR(&rec), cv_fixed_weight(alpha), RecRS(&rec), std::mt19937(std::size_t(seed)), (unsigned long)spp, std::size_t(16)));
        viltrum::integrate(integrator, acc, r, f, range, logger);
        return 0;
    });
}

extern "C" int vo_cv_policies(const char* integrand, uint64_t iterations, uint64_t spp, uint64_t seed, int rr_policy, int weight_strategy, double alpha,
                   int dimbins, const uint64_t* res, const float* rmin, const float* rmax, float* bins,
                   uint32_t* rec_nregions, float* rec_approx, uint32_t* rec_chosen, float* rec_samples) {
    if (rr_policy<0 || rr_policy>3 || weight_strategy<0 || weight_strategy>1) return -3;
    RegionSink sink;
    return dispatch_finite_bins(integrand, dimbins, [&] (auto f, auto dbc) -> int {
        using F = decltype(f);
        constexpr std::size_t D = F::dim;
        constexpr std::size_t DB = decltype(dbc)::value;
        using namespace viltrum;
        auto r = res_array<DB>(res);
        auto range = range_array<D>(rmin, rmax);
        auto acc = [&] (const std::array<std::size_t,DB>& p) -> float& { return bins[tensor_pos(p,r)]; };
        CvRecorder rec; rec.db = int(DB); rec.spp = spp; rec.dim = D; rec.sink = &sink;
        for (std::size_t i=0;i<DB;++i) { rec.rmin[i]=rmin[i]; rec.res[i]=r[i]; rec.drange[i]=(rmax[i]-rmin[i])/float(r[i]); }
        rec.nregions=rec_nregions; rec.approx=nullptr; rec.chosen=rec_chosen; rec.samples=rec_samples;
        DumpLogger logger(&sink);
        auto run = [&] (auto rr, auto cv) {
            auto integrator = integrator_region_based(
                regions_generator_adaptive_heap(nested(simpson,trapezoidal), error_heuristic_size(error_metric_relative(),1.e-5), std::size_t(iterations)),
                regions_integrator_parallel_variance_reduction(std::move(rr), std::move(cv), RecRS(&rec), std::mt19937(std::size_t(seed)), (unsigned long)spp, std::size_t(16)));
            viltrum::integrate(integrator, acc, r, f, range, logger);
        };
        auto with_cv = [&] (auto rr) {
            if (weight_strategy == 1) run(std::move(rr), cv_fixed_weight(alpha));
            else { rec.approx = rec_approx; run(std::move(rr), RecCV(&rec)); }
        };
        if (rr_policy == 0) with_cv(RecRRT<rr_uniform_region>(&rec));
        else if (rr_policy == 1) with_cv(RecRRT<rr_integral_region<>>(&rec));
        else if (rr_policy == 2) with_cv(RecRRT<rr_error_region<>>(&rec));
        else with_cv(RecRRT<rr_pdf_region<>>(&rec));
        return 0;
    });
}


// The crespo2021 pipeline with the region-SAMPLING policy chosen by the caller (reference src/control-variates/region-sampling.h:9-135):
// rs_policy 0 region_sampling_uniform, 1 region_sampling_importance, 2 region_sampling_mis(power, cutoff), 3 region_sampling_russian_roulette;
// rr_uniform_region, cv_optimize_weight.  No recording (the importance samplers draw through Simpson::sample: statistical parity only).
extern "C" int vo_cv_sampling(const char* integrand, uint64_t iterations, uint64_t spp, uint64_t seed, int rs_policy, double power, double cutoff,
                   int dimbins, const uint64_t* res, const float* rmin, const float* rmax, float* bins) {
    if (rs_policy < 0 || rs_policy > 3) return -3;
    return dispatch_finite_bins(integrand, dimbins, [&] (auto f, auto dbc) -> int {
        using F = decltype(f);
        constexpr std::size_t D = F::dim;
        constexpr std::size_t DB = decltype(dbc)::value;
        using namespace viltrum;
        auto r = res_array<DB>(res);
        auto range = range_array<D>(rmin, rmax);
        auto acc = [&] (const std::array<std::size_t,DB>& p) -> float& { return bins[tensor_pos(p,r)]; };
        auto run = [&] (auto rs) {
            auto integrator = integrator_region_based(
                regions_generator_adaptive_heap(nested(simpson,trapezoidal), error_heuristic_size(error_metric_relative(),1.e-5), std::size_t(iterations)),
                regions_integrator_parallel_variance_reduction(rr_uniform_region(), cv_optimize_weight(), std::move(rs), std::mt19937(std::size_t(seed)), (unsigned long)spp, std::size_t(16)));
            viltrum::integrate(integrator, acc, r, f, range);
        };
        if (rs_policy == 0) run(region_sampling_uniform());
        else if (rs_policy == 1) run(region_sampling_importance<>());
        else if (rs_policy == 2) run(region_sampling_mis<>(power, cutoff));
        else run(region_sampling_russian_roulette<>());
        return 0;
    });
}
