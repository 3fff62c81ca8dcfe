// viltrum_b200/viltrum.h — the reference's C++17 vocabulary for the per-bin integration hot path, on the GPU.
//
// Drop-in for `#include "viltrum.h"` of adolfomunoz/viltrum as far as the hot path goes: the same
// namespace, free function and factories —
//     viltrum::integrate(integrator, bins, resolution, integrand, range[, logger])        reference src/integrate.h:72-103
//     viltrum::integrate(integrator, std::vector<T>& bins, integrand, range[, logger])    reference src/integrate.h:132-137,169-173
//     monte_carlo, monte_carlo_per_bin_parallel, integrator_per_bin_parallel, integrator_newton_cotes,
//     integrator_adaptive_iterations, integrator_adaptive_tolerance, integrator_crespo2021, integrator_fubini<N>, integrator_crespo2021_infinite<N>,
//     integrator_adaptive_variance_reduction_parallel with rr_uniform_region / rr_integral_region / rr_error_region / rr_pdf_region,
//     cv_optimize_weight / cv_fixed_weight, region_sampling_uniform,
//     range_split_at<N>, nested, trapezoidal / simpson / boole,
//     error_heuristic_default / error_heuristic_size, error_metric_absolute / error_metric_relative,
//     range, range_all, range_primary, range_infinite, range_primary_infinite, tensor, LoggerNull, LoggerProgress
// — with every integrator's integrate(bins, resolution, f, range, logger) member forwarding to libviltrum_b200.so through
// the C ABI of include/viltrum_b200.h.  What changes for the caller:
//   * the translation unit is compiled by nvcc (-std=c++17 --expt-relaxed-constexpr [--extended-lambda]
//     -gencode arch=compute_100a,code=sm_100a) and linked with -lviltrum_b200;
//   * integrands are __device__-callable, trivially copyable functors (or __host__ __device__ extended lambdas):
//       finite   : float operator()(const std::array<float,DIM>&) const   — or 1..7 float arguments (integrate.h:13-69)
//       infinite : template<class Seq> float operator()(const Seq&) const   (seq.begin(), *it, ++it — doc/integrands.md:112-143)
//   * Float is float and the bin value type is float (the reference also allows double and vector-valued bins);
//   * random numbers come from Philox4x32-10 keyed by (seed; bin, sample), not from std::mt19937: results agree with the
//     reference statistically (and bit for bit in the replay / exact modes of the C ABI), not stream for stream.
// Bin write semantics follow the reference integrator by integrator ('+=' / '=').  There is no CPU
// fallback: without a CUDA device the first integrate() throws std::runtime_error.
#pragma once
#include <array>
#include <vector>
#include <tuple>
#include <string>
#include <stdexcept>
#include <limits>
#include <chrono>
#include <iostream>
#include <iomanip>
#include <type_traits>
#include <cstring>
#include <cmath>
#include "../viltrum_b200.h"
#include "device/thunks.cuh"
#include "device/fubini.cuh"

namespace viltrum {

// ---- ranges (reference src/range.h:17-199, src/range-infinite.h:16-132) ----------------------------------------------
template<typename T, std::size_t DIM>
class Range : public std::array<std::array<T,DIM>,2> {
    T _volume;
public:
    static constexpr std::size_t dimensions = DIM;
    using value_type = T;
    static constexpr std::size_t size = DIM;
    Range(const std::array<T,DIM>& a, const std::array<T,DIM>& b) : std::array<std::array<T,DIM>,2>{a,b} {
        _volume = T(1); for (std::size_t i = 0; i<DIM; ++i) _volume*=(b[i]-a[i]);
    }
    const std::array<T,DIM>& min() const { return (*this)[0]; }
    T min(std::size_t i) const { return min()[i]; }
    const std::array<T,DIM>& max() const { return (*this)[1]; }
    T max(std::size_t i) const { return max()[i]; }
    T volume() const { return _volume; }
    bool is_inside(const std::array<T,DIM>& x) const { bool is = true; for (std::size_t i = 0; (i<DIM) && is; ++i) is = ((x[i]>=min(i)) && (x[i]<=max(i))); return is; }
    Range<T,DIM> subrange_dimension(std::size_t dim, T a, T b) const { auto na = min(); na[dim]=a; auto nb = max(); nb[dim]=b; return Range<T,DIM>(na,nb); }
    bool empty() const { bool e = false; for (std::size_t i = 0; (i<DIM) && (!e); ++i) e = (min(i)>=max(i)); return e; }
};
template<typename T, std::size_t DIM> Range<T,DIM> range(const std::array<T,DIM>& a, const std::array<T,DIM>& b) { return Range<T,DIM>(a,b); }
template<typename T> Range<T,1> range(const T& a, const T& b, std::enable_if_t<std::is_floating_point_v<T>,int> = 0) { return Range<T,1>(std::array<T,1>{a},std::array<T,1>{b}); }
template<typename T> Range<T,2> range(const T& a0, const T& a1, const T& b0, const T& b1) { return Range<T,2>(std::array<T,2>{a0,a1},std::array<T,2>{b0,b1}); }
template<typename T> Range<T,3> range(const T& a0, const T& a1, const T& a2, const T& b0, const T& b1, const T& b2) { return Range<T,3>(std::array<T,3>{a0,a1,a2},std::array<T,3>{b0,b1,b2}); }
template<std::size_t N, typename T> Range<T,N> range_all(const T& va, const T& vb) { std::array<T,N> a, b; a.fill(va); b.fill(vb); return range(a,b); }
template<std::size_t N, typename T = float> Range<T,N> range_primary() { return range_all<N>(T(0),T(1)); }

template<typename T>
class RangeInfinite : public std::array<std::vector<T>,2> {
    T _volume;
public:
    static constexpr std::size_t dimensions = std::numeric_limits<std::size_t>::max();
    using value_type = T;
    RangeInfinite(const std::vector<T>& a = std::vector<T>(), const std::vector<T>& b = std::vector<T>()) : std::array<std::vector<T>,2>{a,b} {
        _volume = T(1); for (std::size_t i = 0; i<std::max(a.size(),b.size()); ++i) _volume*=(max(i)-min(i));
    }
    const std::vector<T>& min() const { return (*this)[0]; }
    T min(std::size_t i) const { return (i<min().size())?(min()[i]):T(0); }
    const std::vector<T>& max() const { return (*this)[1]; }
    T max(std::size_t i) const { return (i<max().size())?(max()[i]):T(1); }
    T volume() const { return _volume; }
};
template<typename T> RangeInfinite<T> range_infinite(const std::vector<T>& a, const std::vector<T>& b) { return RangeInfinite<T>(a,b); }
template<typename T> RangeInfinite<T> range_infinite(const T& a, const T& b, std::enable_if_t<std::is_floating_point_v<T>,int> = 0) { return RangeInfinite<T>(std::vector<T>{a},std::vector<T>{b}); }
template<typename T = float> RangeInfinite<T> range_primary_infinite() { return RangeInfinite<T>(); }

// ---- tensor (reference src/tensor.h:8-57): flat storage, dimension 0 fastest ----------------------------------------------
template<typename T, std::size_t DIMBINS>
class tensor {
    std::vector<T> data_; std::array<std::size_t, DIMBINS> res;
    std::size_t position(const std::array<std::size_t,DIMBINS>& p) const { std::size_t pos = 0, prod = 1; for (std::size_t d = 0;d<DIMBINS; ++d) { pos += p[d]*prod; prod*=res[d]; } return pos; }
public:
    tensor(const std::array<std::size_t, DIMBINS>& r, const T& t = T()) : res(r) { std::size_t n(1); for (auto x : res) n*=x; data_.resize(n, t); }
    const std::array<std::size_t, DIMBINS>& resolution() const { return res; }
    std::size_t resolution(std::size_t i) const { return res[i]; }
    const std::vector<T>& raw_data() const { return data_; }
    T* data() { return data_.data(); }                       // (addition) lets the GPU back end skip the per-bin accessor loop
    T& operator[](const std::array<std::size_t,DIMBINS>& p) { return data_[position(p)]; }
    T& operator()(const std::array<std::size_t,DIMBINS>& p) { return data_[position(p)]; }
    const T& operator[](const std::array<std::size_t,DIMBINS>& p) const { return data_[position(p)]; }
    const T& operator()(const std::array<std::size_t,DIMBINS>& p) const { return data_[position(p)]; }
    std::size_t size() const { return data_.size(); }
};

// ---- loggers (reference src/log.h:10-52) ----------------------------------------------------------------------------------
class LoggerNull {
public:
    LoggerNull(const std::string& = "") {}
    std::string name() const { return ""; }
    void set_name(const std::string&) {}
    template<typename Number> void log_progress(const Number&, const Number& = Number(1)) {}
    template<typename Data> void log(const Data&) {}
};
class LoggerProgress {
    std::string name_; std::chrono::time_point<std::chrono::steady_clock> start = std::chrono::steady_clock::now();
public:
    LoggerProgress(const std::string& n) : name_(n) {}
    const std::string& name() const { return name_; }
    void set_name(const std::string& name) { name_=name; }
    template<typename Number> void log_progress(const Number& number, const Number& last = Number(1)) {
        if (number<=Number(0)) { start = std::chrono::steady_clock::now(); std::cerr<<name()<<" -                           \r"; }
        else if (number >= last) { auto el = std::chrono::duration_cast<std::chrono::duration<double>>(std::chrono::steady_clock::now() - start);
            std::cerr<<name()<<" - \t [DONE]\t("<<std::setprecision(3)<<std::setw(6)<<el.count()<<" seconds)\n"; }
    }
    template<typename Data> void log(const Data&) {}
};
template<typename Logger> Logger logger_step(const Logger& logger, std::string step_name) { Logger sol = logger; sol.set_name(logger.name()+" | "+step_name); return sol; }

// ---- GPU plumbing -----------------------------------------------------------------------------------------------------------
namespace b200 {

class Context {
    vb200_ctx* h = nullptr;
public:
    explicit Context(int device = 0) { if (vb200_create(device, &h) != VB200_OK) throw std::runtime_error(std::string("viltrum_b200: ") + vb200_last_error(nullptr)); }
    ~Context() { vb200_destroy(h); }
    Context(const Context&) = delete; Context& operator=(const Context&) = delete;
    vb200_ctx* get() const { return h; }
    void check(int status) const { if (status != VB200_OK) throw std::runtime_error(std::string("viltrum_b200: ") + vb200_last_error(h)); }
};
// one context per process and device, created on first use (one process per GPU)
inline int& current_device() { static int d = 0; return d; }
inline Context& default_context() { static Context c(current_device()); return c; }
// (addition) pin + map a tensor's storage so that the per-bin samplers write their bins into it directly over PCIe, without the
// staged copy and the host-side '+=' pass (vb200_host_register).  Unpin before the tensor is destroyed or resized.
template<typename T, std::size_t DIMBINS> inline void pin_bins(tensor<T,DIMBINS>& t) {
    auto& ctx = default_context(); ctx.check(vb200_host_register(ctx.get(), t.data(), t.size()*sizeof(T)));
}
template<typename T, std::size_t DIMBINS> inline void unpin_bins(tensor<T,DIMBINS>& t) {
    auto& ctx = default_context(); ctx.check(vb200_host_unregister(ctx.get(), t.data()));
}
// bin-grid shard handled by this process ({0,0} = whole grid): multi-GPU runs set it per rank
inline vb200_shard& current_shard() { static vb200_shard s{0, 0}; return s; }
// slab of `rank` out of `world`: whole rows of the LAST bin dimension, one contiguous range of linear bin indices; a rank beyond the
// number of rows gets the explicit empty shard (VB200_SHARD_EMPTY_INDEX), never {0,0} (= whole grid)
template<std::size_t DIMBINS>
inline vb200_shard shard_for_rank(const std::array<std::size_t,DIMBINS>& res, std::size_t rank, std::size_t world) {
    std::size_t stride = 1; for (std::size_t i = 0; i + 1 < DIMBINS; ++i) stride *= res[i];
    const std::size_t rows = res[DIMBINS-1], lo = rows * rank / world, hi = rows * (rank + 1) / world;
    if (lo == hi) return vb200_shard{VB200_SHARD_EMPTY_INDEX, VB200_SHARD_EMPTY_INDEX};
    return vb200_shard{uint64_t(lo) * stride, uint64_t(hi) * stride};
}
// process-wide sampler options of the per-bin Monte-Carlo integrators (vb200_mc_params.options, VB200_MC_* flags): 0 = xoshiro128++ streams
// keyed by Philox + 16-bit in-bin lattice on fine grids; VB200_MC_RNG_PHILOX = every draw from Philox4x32-10; VB200_MC_LATTICE24
inline int32_t& mc_options() { static int32_t o = 0; return o; }

template<typename Float, std::size_t DIM, std::size_t DIMBINS>
inline vb200_domain make_domain(const Range<Float,DIM>& r, const std::array<std::size_t,DIMBINS>& res) {
    static_assert(std::is_same<Float,float>::value, "viltrum_b200 computes in fp32: use Range<float,DIM>");
    static_assert(DIM <= VB200_MAX_DIM && DIMBINS <= VB200_MAX_DIMBINS && DIMBINS <= DIM, "unsupported dimensionality");
    vb200_domain d; std::memset(&d, 0, sizeof(d));
    d.dim = int(DIM); d.dimbins = int(DIMBINS);
    for (std::size_t i = 0; i < DIM; ++i) { d.rmin[i] = r.min(i); d.rmax[i] = r.max(i); }
    for (std::size_t i = 0; i < DIMBINS; ++i) d.res[i] = res[i];
    return d;
}
template<typename Float, std::size_t DIMBINS>
inline vb200_domain make_domain(const RangeInfinite<Float>& r, const std::array<std::size_t,DIMBINS>& res) {
    static_assert(std::is_same<Float,float>::value, "viltrum_b200 computes in fp32: use RangeInfinite<float>");
    vb200_domain d; std::memset(&d, 0, sizeof(d));
    const std::size_t n = std::max(r.min().size(), r.max().size());
    if (n > VB200_MAX_DIM) throw std::runtime_error("viltrum_b200: at most 8 explicit entries in an infinite range");
    d.dim = int(n); d.dimbins = int(DIMBINS);
    for (std::size_t i = 0; i < n; ++i) { d.rmin[i] = r.min(i); d.rmax[i] = r.max(i); }
    for (std::size_t i = 0; i < DIMBINS; ++i) d.res[i] = res[i];
    return d;
}
template<std::size_t DIM, std::size_t DIMBINS>
inline vb200_domain_f64 make_domain64(const Range<double,DIM>& r, const std::array<std::size_t,DIMBINS>& res) {
    static_assert(DIM <= VB200_MAX_DIM && DIMBINS <= VB200_MAX_DIMBINS && DIMBINS <= DIM, "unsupported dimensionality");
    vb200_domain_f64 d; std::memset(&d, 0, sizeof(d));
    d.dim = int(DIM); d.dimbins = int(DIMBINS);
    for (std::size_t i = 0; i < DIM; ++i) { d.rmin[i] = r.min(i); d.rmax[i] = r.max(i); }
    for (std::size_t i = 0; i < DIMBINS; ++i) d.res[i] = res[i];
    return d;
}
template<std::size_t DIMBINS> inline std::size_t bin_count(const std::array<std::size_t,DIMBINS>& res) { std::size_t n = 1; for (auto r : res) n *= r; return n; }

// applies a flat device result (tensor order) to the caller's bins through the accessor: bins(pos) op= v
template<bool ACCUMULATE, typename Bins, std::size_t DIMBINS, typename V>
inline void apply_bins(Bins& bins, const std::array<std::size_t,DIMBINS>& res, const std::vector<V>& flat) {
    std::array<std::size_t,DIMBINS> pos; pos.fill(0);
    vb200_shard sh = current_shard(); const std::size_t n = flat.size();
    if (sh.begin == VB200_SHARD_EMPTY_INDEX && sh.end == VB200_SHARD_EMPTY_INDEX) return;      // explicit empty shard
    const std::size_t b = (sh.begin == 0 && sh.end == 0) ? 0 : std::size_t(sh.begin), e = (sh.begin == 0 && sh.end == 0) ? n : std::size_t(sh.end);
    for (std::size_t k = 0; k < n; ++k) {
        if (k >= b && k < e) { if (ACCUMULATE) bins(pos) += flat[k]; else bins(pos) = flat[k]; }
        for (std::size_t d = 0; d < DIMBINS; ++d) { if (++pos[d] >= res[d]) pos[d] = 0; else break; }
    }
}

// scalar-argument integrands f(x0, x1, ...) -> array integrands (reference src/integrate.h:13-69), device-callable
template<typename F, std::size_t N> struct ScalarAdapter {
    F f;
    template<std::size_t... I> __host__ __device__ float call(const std::array<float,N>& x, std::index_sequence<I...>) const { return f(x[I]...); }
    __host__ __device__ float operator()(const std::array<float,N>& x) const { return call(x, std::make_index_sequence<N>()); }
};
template<typename F, std::size_t N, typename = void> struct takes_array : std::false_type {};
template<typename F, std::size_t N> struct takes_array<F, N, std::void_t<decltype(std::declval<const F&>()(std::declval<const std::array<float,N>&>()))>> : std::true_type {};
template<typename F, std::size_t N, typename = void> struct takes_array_d : std::false_type {};
template<typename F, std::size_t N> struct takes_array_d<F, N, std::void_t<decltype(std::declval<const F&>()(std::declval<const std::array<double,N>&>()))>> : std::true_type {};
template<std::size_t DIM, typename F>
inline auto adapt(const F& f) {
    if constexpr (takes_array<F, DIM>::value || takes_array_d<F, DIM>::value) return f;
    else return ScalarAdapter<F, DIM>{f};
}

// region list handed to Logger::log (the reference passes its vector of regions, integrator-region-based.h:19)
template<std::size_t DIM>
struct RegionView {
    Range<float,DIM> box; std::tuple<float,std::size_t> errdim; std::vector<float> samples;
    const Range<float,DIM>& range() const { return box; }
    const std::tuple<float,std::size_t>& extra() const { return errdim; }
};
template<std::size_t DIM>
inline std::vector<RegionView<DIM>> download_regions(Context& ctx, const vb200_regions* r) {
    const std::size_t n = vb200_regions_count(r), sd = std::size_t(vb200_regions_samples(r));
    std::vector<float> mn(n*DIM), mx(n*DIM), err(n), data(n*sd); std::vector<uint32_t> dim(n);
    ctx.check(vb200_regions_download(ctx.get(), r, mn.data(), mx.data(), err.data(), dim.data(), data.data()));
    std::vector<RegionView<DIM>> out; out.reserve(n);
    for (std::size_t i = 0; i < n; ++i) {
        std::array<float,DIM> a, b; for (std::size_t d = 0; d < DIM; ++d) { a[d] = mn[i*DIM+d]; b[d] = mx[i*DIM+d]; }
        out.push_back(RegionView<DIM>{Range<float,DIM>(a,b), std::tuple<float,std::size_t>(err[i], dim[i]), std::vector<float>(data.begin()+i*sd, data.begin()+(i+1)*sd)});
    }
    return out;
}
struct RegionsHandle { vb200_regions* r = nullptr; ~RegionsHandle() { vb200_regions_free(r); } };
// the one-bin accessor of the value-returning integrate overloads (integrate.h:105-123)
struct SingleBin { float* value; float& operator()(const std::array<std::size_t,1>&) const { return *value; } };

} // namespace b200

// ---- Monte Carlo integrators -----------------------------------------------------------------------------------------------
// monte_carlo(samples, seed) — reference src/monte-carlo/monte-carlo.h:39-63,88-95: global sampler, bins(pos) += f*factor
class MonteCarlo {
    unsigned long samples; std::size_t seed_;
public:
    MonteCarlo(unsigned long s, std::size_t seed) : samples(s), seed_(seed) {}
    unsigned long sample_count() const { return samples; }
    std::size_t seed() const { return seed_; }
    template<typename Bins, std::size_t DIMBINS, typename F, typename Float, std::size_t DIM, typename Logger>
    void integrate(Bins& bins, const std::array<std::size_t,DIMBINS>& res, const F& f, const Range<Float,DIM>& range, Logger& logger) const {
        auto& ctx = b200::default_context();
        b200::Integrand<F, int(DIM)> g(f);
        vb200_mc_params p; std::memset(&p, 0, sizeof(p));
        p.domain = b200::make_domain(range, res); p.spp = samples; p.seed = seed_; p.flavor = VB200_MC_PER_BIN;
        std::vector<float> flat(b200::bin_count(res), 0.0f);
        logger.log_progress(0ul, samples);
        ctx.check(vb200_monte_carlo(ctx.get(), g.c_abi(), &p, flat.data(), VB200_HOST));
        b200::apply_bins<true>(bins, res, flat);
        logger.log_progress(samples, samples);
    }
    // global sampler over an infinite range (monte-carlo.h:65-84): the bin comes from the first DIMBINS sequence elements
    template<typename Bins, std::size_t DIMBINS, typename F, typename Float, typename Logger>
    void integrate(Bins& bins, const std::array<std::size_t,DIMBINS>& res, const F& f, const RangeInfinite<Float>& range, Logger& logger) const {
        auto& ctx = b200::default_context();
        b200::InfiniteIntegrand<F> g(f);
        vb200_mc_params p; std::memset(&p, 0, sizeof(p));
        p.domain = b200::make_domain(range, res); p.spp = samples; p.seed = seed_; p.flavor = VB200_MC_PER_BIN;
        std::vector<float> flat(b200::bin_count(res), 0.0f);
        ctx.check(vb200_monte_carlo(ctx.get(), g.c_abi(), &p, flat.data(), VB200_HOST));
        b200::apply_bins<true>(bins, res, flat);
        logger.log_progress(samples, samples);
    }
};
inline MonteCarlo monte_carlo(unsigned long samples, std::size_t seed = 0) { return MonteCarlo(samples, seed); }

// monte_carlo_per_bin_parallel(spp, seed) — reference src/monte-carlo/monte-carlo-per-bin-parallel.h:41-100,104-111 ('+=')
class MonteCarloPerBinParallel {
    unsigned long samples; std::size_t seed_;
    template<bool INF, typename Bins, std::size_t DIMBINS, typename G, typename Logger>
    void run(Bins& bins, const std::array<std::size_t,DIMBINS>& res, const G& g, const vb200_domain& dom, Logger& logger) const {
        auto& ctx = b200::default_context();
        vb200_mc_params p; std::memset(&p, 0, sizeof(p));
        p.domain = dom; p.shard = b200::current_shard(); p.spp = samples; p.seed = seed_; p.flavor = VB200_MC_PER_BIN; p.options = b200::mc_options();
        logger.log_progress(std::size_t(0), std::size_t(1));
        if constexpr (std::is_same<Bins, tensor<float,DIMBINS>>::value) {      // fast path: the library accumulates straight into the tensor
            ctx.check(INF ? vb200_mc_per_bin_inf(ctx.get(), g.c_abi(), &p, bins.data(), VB200_HOST, nullptr, nullptr)
                          : vb200_mc_per_bin(ctx.get(), g.c_abi(), &p, bins.data(), VB200_HOST, nullptr, nullptr));
        } else {
            std::vector<float> flat(b200::bin_count(res), 0.0f);
            ctx.check(INF ? vb200_mc_per_bin_inf(ctx.get(), g.c_abi(), &p, flat.data(), VB200_HOST, nullptr, nullptr)
                          : vb200_mc_per_bin(ctx.get(), g.c_abi(), &p, flat.data(), VB200_HOST, nullptr, nullptr));
            b200::apply_bins<true>(bins, res, flat);
        }
        logger.log_progress(std::size_t(1), std::size_t(1));
    }
public:
    MonteCarloPerBinParallel(unsigned long s, std::size_t seed) : samples(s), seed_(seed) {}
    template<typename Bins, std::size_t DIMBINS, typename F, typename Float, std::size_t DIM, typename Logger>
    void integrate(Bins& bins, const std::array<std::size_t,DIMBINS>& res, const F& f, const Range<Float,DIM>& range, Logger& logger) const {
        b200::Integrand<F, int(DIM)> g(f);
        run<false>(bins, res, g, b200::make_domain(range, res), logger);
    }
    template<typename Bins, std::size_t DIMBINS, typename F, typename Float, typename Logger>
    void integrate(Bins& bins, const std::array<std::size_t,DIMBINS>& res, const F& f, const RangeInfinite<Float>& range, Logger& logger) const {
        b200::InfiniteIntegrand<F> g(f);
        run<true>(bins, res, g, b200::make_domain(range, res), logger);
    }
};
inline MonteCarloPerBinParallel monte_carlo_per_bin_parallel(unsigned long samples, std::size_t seed = 0) { return MonteCarloPerBinParallel(samples, seed); }
inline MonteCarloPerBinParallel monte_carlo_per_bin(unsigned long samples, std::size_t seed = 0) { return MonteCarloPerBinParallel(samples, seed); }   // monte-carlo-per-bin.h: same estimator, sequential upstream

// integrator_per_bin_parallel(monte_carlo(spp, seed)) — reference src/integrator-per-bin-parallel.h:16-38 ('=')
template<typename Integrator> class IntegratorPerBinParallel;
template<> class IntegratorPerBinParallel<MonteCarlo> {
    MonteCarlo inner;
public:
    IntegratorPerBinParallel(const MonteCarlo& m) : inner(m) {}
    template<typename Bins, std::size_t DIMBINS, typename F, typename Float, std::size_t DIM, typename Logger>
    void integrate(Bins& bins, const std::array<std::size_t,DIMBINS>& res, const F& f, const Range<Float,DIM>& range, Logger& logger) const {
        auto& ctx = b200::default_context();
        b200::Integrand<F, int(DIM)> g(f);
        vb200_mc_params p; std::memset(&p, 0, sizeof(p));
        p.domain = b200::make_domain(range, res); p.shard = b200::current_shard(); p.spp = inner.sample_count(); p.seed = inner.seed(); p.flavor = VB200_PER_BIN_MC; p.options = b200::mc_options();
        std::vector<float> flat(b200::bin_count(res), 0.0f);
        logger.log_progress(std::size_t(0), std::size_t(1));
        ctx.check(vb200_mc_per_bin(ctx.get(), g.c_abi(), &p, flat.data(), VB200_HOST, nullptr, nullptr));
        b200::apply_bins<false>(bins, res, flat);
        logger.log_progress(std::size_t(1), std::size_t(1));
    }
    // row a20: the wrapper spelling over an infinite range (monte-carlo.h:65-84 per bin), '='
    template<typename Bins, std::size_t DIMBINS, typename F, typename Float, typename Logger>
    void integrate(Bins& bins, const std::array<std::size_t,DIMBINS>& res, const F& f, const RangeInfinite<Float>& range, Logger& logger) const {
        auto& ctx = b200::default_context();
        b200::InfiniteIntegrand<F> g(f);
        vb200_mc_params p; std::memset(&p, 0, sizeof(p));
        p.domain = b200::make_domain(range, res); p.shard = b200::current_shard(); p.spp = inner.sample_count(); p.seed = inner.seed(); p.flavor = VB200_PER_BIN_MC;
        std::vector<float> flat(b200::bin_count(res), 0.0f);
        ctx.check(vb200_mc_per_bin_inf(ctx.get(), g.c_abi(), &p, flat.data(), VB200_HOST, nullptr, nullptr));
        b200::apply_bins<false>(bins, res, flat);
        logger.log_progress(std::size_t(1), std::size_t(1));
    }
};
inline IntegratorPerBinParallel<MonteCarlo> integrator_per_bin_parallel(const MonteCarlo& m) { return IntegratorPerBinParallel<MonteCarlo>(m); }
inline IntegratorPerBinParallel<MonteCarlo> integrator_per_bin(const MonteCarlo& m) { return IntegratorPerBinParallel<MonteCarlo>(m); }

// ---- Newton-Cotes rules, nested pairs, error heuristics (reference src/newton-cotes/rules.h, src/nested/*.h) ------------------
struct Trapezoidal { static constexpr std::size_t samples = 2; static constexpr int id = VB200_RULE_TRAPEZOIDAL; };
struct Simpson { static constexpr std::size_t samples = 3; static constexpr int id = VB200_RULE_SIMPSON; };
struct Boole { static constexpr std::size_t samples = 5; static constexpr int id = VB200_RULE_BOOLE; };
static const Trapezoidal trapezoidal{}; static const Simpson simpson{}; static const Boole boole{};
// steps<N>(rule) — reference rules.h:321-389: N pieces of rule Q per dimension (fixed-rule integration)
template<typename Q, std::size_t N> struct Steps { static constexpr std::size_t samples = (Q::samples - 1)*N + 1; static constexpr int id = VB200_RULE_STEPS(int(Q::samples), int(N)); };
template<std::size_t N, typename Q> Steps<Q,N> steps(const Q&) { static_assert(N >= 1 && N <= 65535, "steps<N>: 1 <= N <= 65535"); return Steps<Q,N>(); }
template<typename H, typename L> struct Nested {
    static constexpr std::size_t samples = H::samples;
    static_assert((std::is_same<H,Simpson>::value && std::is_same<L,Trapezoidal>::value) || (std::is_same<H,Boole>::value && std::is_same<L,Simpson>::value),
                  "viltrum_b200 implements nested(simpson,trapezoidal) and nested(boole,simpson)");
    static constexpr int id = std::is_same<H,Simpson>::value ? VB200_RULE_SIMPSON_TRAPEZOIDAL : VB200_RULE_BOOLE_SIMPSON;
};
template<typename H, typename L> Nested<H,L> nested(const H&, const L&) { return Nested<H,L>(); }
struct error_metric_absolute { static constexpr int id = VB200_METRIC_ABSOLUTE; };
struct error_metric_relative { static constexpr int id = VB200_METRIC_RELATIVE; error_metric_relative(double = 1.e-37) {} };
template<typename EM> struct error_heuristic_default { static constexpr int id = VB200_HEURISTIC_DEFAULT; double size_weight = 0; error_heuristic_default(const EM&) {} using metric = EM;
    void fill(vb200_mixed_heuristic&) const {} };
template<typename EM> struct error_heuristic_size { static constexpr int id = VB200_HEURISTIC_SIZE; double size_weight; error_heuristic_size(const EM&, double sw = 1.e-5,
        double = 1.e-37) : size_weight(sw) {} using metric = EM;
    void fill(vb200_mixed_heuristic&) const {} };
// error_heuristic_mixed(metric_bins, metric_rest, dimension, bins_weight, size_weight, size_bins, size_rest, error_increase_factor) — reference
// src/nested/error-heuristic.h:49-98, same argument order and defaults
template<typename EMB, typename EMR> struct error_heuristic_mixed {
    static constexpr int id = VB200_HEURISTIC_MIXED; using metric = EMB;
    double size_weight; vb200_mixed_heuristic m;
    error_heuristic_mixed(const EMB&, const EMR&, unsigned int dim = 2, double bins_w = 1.0, double sw = 1.e-3, double size_bins = 1.0/1024.0, double size_rest = 1.0/16.0,
                          double error_increase_factor = 1.e4) : size_weight(sw), m{EMR::id, int(dim), bins_w, size_bins, size_rest, error_increase_factor} {}
    void fill(vb200_mixed_heuristic& out) const { out = m; }
};

// integrator_newton_cotes(rule) — reference src/newton-cotes/newton-cotes.h:11-19 ('+=')
template<typename Rule> class IntegratorNewtonCotes {
public:
    // Range<double,DIM> with a double integrand: the fp64 path (every rule and fold in double, as upstream)
    template<typename Bins, std::size_t DIMBINS, typename F, std::size_t DIM, typename Logger>
    void integrate(Bins& bins, const std::array<std::size_t,DIMBINS>& res, const F& f, const Range<double,DIM>& range, Logger& logger) const {
        auto& ctx = b200::default_context();
        b200::Integrand64<F, int(DIM)> g(f);
        vb200_domain_f64 dom = b200::make_domain64(range, res);
        b200::RegionsHandle regs;
        ctx.check(vb200_regions_generate_single_f64(ctx.get(), g.c_abi(), &dom, Rule::id, &regs.r));
        std::vector<double> flat(b200::bin_count(res), 0.0);
        vb200_shard sh = b200::current_shard();
        ctx.check(vb200_regions_integrate_bins_f64(ctx.get(), regs.r, &dom, &sh, flat.data(), VB200_HOST));
        b200::apply_bins<true>(bins, res, flat);
        logger.log_progress(std::size_t(1), std::size_t(1));
    }
    template<typename Bins, std::size_t DIMBINS, typename F, std::size_t DIM, typename Logger>
    void integrate(Bins& bins, const std::array<std::size_t,DIMBINS>& res, const F& f, const Range<float,DIM>& range, Logger& logger) const {
        using Float = float;
        auto& ctx = b200::default_context();
        b200::Integrand<F, int(DIM)> g(f);
        vb200_domain dom = b200::make_domain(range, res);
        b200::RegionsHandle regs;
        ctx.check(vb200_regions_generate_single(ctx.get(), g.c_abi(), &dom, Rule::id, &regs.r));
        std::vector<float> flat(b200::bin_count(res), 0.0f);
        vb200_shard sh = b200::current_shard();
        ctx.check(vb200_regions_integrate_bins(ctx.get(), regs.r, &dom, &sh, flat.data(), VB200_HOST));
        b200::apply_bins<true>(bins, res, flat);
        logger.log_progress(std::size_t(1), std::size_t(1));
    }
};
template<typename R> IntegratorNewtonCotes<R> integrator_newton_cotes(const R&) { return IntegratorNewtonCotes<R>(); }
template<typename R> IntegratorNewtonCotes<R> integrator_newton_cotes_parallel(const R&) { return IntegratorNewtonCotes<R>(); }

// integrator_adaptive_iterations(nested(h,l), error_heuristic, iterations) — reference src/nested/integrator-adaptive-iterations.h:12-30,
// src/nested/regions-generator-adaptive-heap.h:18-45 + src/newton-cotes/regions-integrator-sequential.h:38-58 ('+=').
// batch = 1 reproduces the reference's greedy split order exactly.
template<typename Rule, typename EH> class IntegratorAdaptiveIterations {
    EH eh; std::size_t iterations; int batch;
public:
    IntegratorAdaptiveIterations(const EH& e, std::size_t it, int b = 1) : eh(e), iterations(it), batch(b) {}
    template<typename F, typename Float, std::size_t DIM>
    void generate(b200::Context& ctx, const b200::Integrand<F,int(DIM)>& g, const Range<Float,DIM>& range, b200::RegionsHandle& regs) const {
        vb200_adaptive_params p; std::memset(&p, 0, sizeof(p));
        std::array<std::size_t,1> one{1}; p.domain = b200::make_domain(range, one);
        p.rule = Rule::id; p.heuristic = EH::id; p.metric = EH::metric::id; p.batch = batch; p.size_weight = eh.size_weight; p.iterations = iterations;
        eh.fill(p.mixed);
        ctx.check(vb200_regions_generate_adaptive(ctx.get(), g.c_abi(), &p, &regs.r));
    }
    // Range<double,DIM> with a double integrand: the greedy generator and the region->bin accumulation in double, as upstream for Float = double
    template<typename Bins, std::size_t DIMBINS, typename F, std::size_t DIM, typename Logger>
    void integrate(Bins& bins, const std::array<std::size_t,DIMBINS>& res, const F& f, const Range<double,DIM>& range, Logger& logger) const {
        auto& ctx = b200::default_context();
        b200::Integrand64<F, int(DIM)> g(f);
        vb200_adaptive_params_f64 p; std::memset(&p, 0, sizeof(p));
        std::array<std::size_t,1> one{1}; p.domain = b200::make_domain64(range, one);
        p.rule = Rule::id; p.heuristic = EH::id; p.metric = EH::metric::id; p.batch = 1; p.size_weight = eh.size_weight; p.iterations = iterations;
        eh.fill(p.mixed);
        b200::RegionsHandle regs;
        ctx.check(vb200_regions_generate_adaptive_f64(ctx.get(), g.c_abi(), &p, &regs.r));
        vb200_domain_f64 dom = b200::make_domain64(range, res);
        std::vector<double> flat(b200::bin_count(res), 0.0);
        vb200_shard sh = b200::current_shard();
        ctx.check(vb200_regions_integrate_bins_f64(ctx.get(), regs.r, &dom, &sh, flat.data(), VB200_HOST));
        b200::apply_bins<true>(bins, res, flat);
        logger.log_progress(iterations, iterations);
    }
    template<typename Bins, std::size_t DIMBINS, typename F, std::size_t DIM, typename Logger>
    void integrate(Bins& bins, const std::array<std::size_t,DIMBINS>& res, const F& f, const Range<float,DIM>& range, Logger& logger) const {
        auto& ctx = b200::default_context();
        b200::Integrand<F, int(DIM)> g(f);
        b200::RegionsHandle regs;
        generate(ctx, g, range, regs);
        if constexpr (!std::is_same<Logger, LoggerNull>::value) logger.log(b200::download_regions<DIM>(ctx, regs.r));
        vb200_domain dom = b200::make_domain(range, res);
        std::vector<float> flat(b200::bin_count(res), 0.0f);
        vb200_shard sh = b200::current_shard();
        ctx.check(vb200_regions_integrate_bins(ctx.get(), regs.r, &dom, &sh, flat.data(), VB200_HOST));
        b200::apply_bins<true>(bins, res, flat);
        logger.log_progress(iterations, iterations);
    }
};
template<typename R, typename EH> auto integrator_adaptive_iterations(const R&, const EH& eh, std::size_t iterations) { return IntegratorAdaptiveIterations<R,EH>(eh, iterations); }
template<typename R, typename EH> auto integrator_adaptive_iterations_parallel(const R&, const EH& eh, std::size_t iterations,
        std::size_t = 16) { return IntegratorAdaptiveIterations<R,EH>(eh, iterations); }
template<typename R> auto integrator_adaptive_iterations(const R& r, std::size_t iterations) { return integrator_adaptive_iterations(r,
        error_heuristic_default<error_metric_absolute>(error_metric_absolute()), iterations); }

// integrator_adaptive_tolerance(nested(h,l), error_heuristic, tolerance) — reference src/nested/integrator-adaptive-tolerance.h:41-59 ('+=').
// Leaves come back in the reference's depth-first order, so the bins match the reference bit for bit in an exact build.
template<typename Rule, typename EH> class IntegratorAdaptiveTolerance {
    EH eh; float tolerance;
public:
    IntegratorAdaptiveTolerance(const EH& e, float tol) : eh(e), tolerance(tol) {}
    template<typename Bins, std::size_t DIMBINS, typename F, typename Float, std::size_t DIM, typename Logger>
    void integrate(Bins& bins, const std::array<std::size_t,DIMBINS>& res, const F& f, const Range<Float,DIM>& range, Logger& logger) const {
        auto& ctx = b200::default_context();
        b200::Integrand<F, int(DIM)> g(f);
        vb200_tolerance_params p; std::memset(&p, 0, sizeof(p));
        std::array<std::size_t,1> one{1}; p.domain = b200::make_domain(range, one);
        p.rule = Rule::id; p.heuristic = EH::id; p.metric = EH::metric::id; p.tolerance = tolerance; p.size_weight = eh.size_weight;
        b200::RegionsHandle regs;
        ctx.check(vb200_regions_generate_tolerance(ctx.get(), g.c_abi(), &p, &regs.r));
        vb200_domain dom = b200::make_domain(range, res);
        std::vector<float> flat(b200::bin_count(res), 0.0f);
        vb200_shard sh = b200::current_shard();
        ctx.check(vb200_regions_integrate_bins(ctx.get(), regs.r, &dom, &sh, flat.data(), VB200_HOST));
        b200::apply_bins<true>(bins, res, flat);
        logger.log_progress(range.volume(), range.volume());
    }
};
template<typename R, typename EH, typename = typename EH::metric> auto integrator_adaptive_tolerance(const R&, const EH& eh,
        float tolerance = 1.e-3f) { return IntegratorAdaptiveTolerance<R,EH>(eh, tolerance); }
template<typename R> auto integrator_adaptive_tolerance(const R& r, float tolerance = 1.e-3f) { return integrator_adaptive_tolerance(r,
        error_heuristic_default<error_metric_absolute>(error_metric_absolute()), tolerance); }
template<typename R> auto integrator_adaptive_tolerance(const R& r, double tolerance) { return integrator_adaptive_tolerance(r, float(tolerance)); }

// integrator_crespo2021(iterations, spp, seed) — reference src/control-variates/integrator-crespo2021.h:7-22 ('=')
// control-variate weight policies (reference src/control-variates/weight-strategy.h:7-110)
struct cv_optimize_weight { static constexpr int id = VB200_CV_OPTIMIZE_WEIGHT; double alpha = 1.0; };
struct cv_fixed_weight { static constexpr int id = VB200_CV_FIXED_WEIGHT; double alpha; cv_fixed_weight(double a = 1) : alpha(a) {} };
// Russian roulette among the regions of a bin (reference src/control-variates/region-russian-roulette.h)
struct rr_uniform_region { static constexpr int id = VB200_RR_UNIFORM; };       // :9-28
struct rr_integral_region { static constexpr int id = VB200_RR_INTEGRAL; };     // :30-67  (NormDefault)
struct rr_error_region { static constexpr int id = VB200_RR_ERROR; };           // :69-106 (NormDefault)
struct rr_pdf_region { static constexpr int id = VB200_RR_PDF; };               // :108-147 (factor_prob = 0.01, NormDefault)
// where inside bin ∩ region a residual sample falls (reference src/control-variates/region-sampling.h:9-135); NormDefault
struct region_sampling_uniform { static constexpr int id = VB200_RS_UNIFORM; double power = 1, cutoff = 0; };                      // :9-20
struct region_sampling_importance { static constexpr int id = VB200_RS_IMPORTANCE; double power = 1, cutoff = 0; };                // :22-44
struct region_sampling_russian_roulette { static constexpr int id = VB200_RS_RUSSIAN_ROULETTE; double power = 1, cutoff = 0; };    // :46-83
struct region_sampling_mis { static constexpr int id = VB200_RS_MIS; double power, cutoff; region_sampling_mis(double p = 1, double c = 0) : power(p), cutoff(c) {} };   // :85-135
// region_stratification_uniform (region-stratification.h:9-25): the sample allocation of the ...Optimized integrators, in the RR slot of their factories
struct region_stratification_uniform { static constexpr int id = VB200_RR_STRATIFIED; };
// integrator_region_based(regions_generator_adaptive_heap(rule, heuristic, iterations), regions_integrator_parallel_variance_reduction(RR, CV,
// region_sampling_uniform, spp, seed)) — reference integrator-adaptive-variance-reduction.h:11-49, regions-integrator-parallel-variance-reduction.h:32-109
template<typename Rule, typename EH> class IntegratorAdaptiveVarianceReduction {
    EH eh; std::size_t iterations, spp, seed_; int weight_strategy = VB200_CV_OPTIMIZE_WEIGHT; double alpha = 1.0; int rr = VB200_RR_UNIFORM;
    int rs = VB200_RS_UNIFORM; double rs_power = 1.0, rs_cutoff = 0.0;
public:
    IntegratorAdaptiveVarianceReduction(const EH& e, std::size_t it, std::size_t s, std::size_t seed, int ws = VB200_CV_OPTIMIZE_WEIGHT, double a = 1.0, int rr_policy = VB200_RR_UNIFORM,
                                        int rs_policy = VB200_RS_UNIFORM, double power = 1.0, double cutoff = 0.0)
        : eh(e), iterations(it), spp(s), seed_(seed), weight_strategy(ws), alpha(a), rr(rr_policy), rs(rs_policy), rs_power(power), rs_cutoff(cutoff) {}
    template<typename Bins, std::size_t DIMBINS, typename F, typename Float, std::size_t DIM, typename Logger>
    void integrate(Bins& bins, const std::array<std::size_t,DIMBINS>& res, const F& f, const Range<Float,DIM>& range, Logger& logger) const {
        auto& ctx = b200::default_context();
        b200::Integrand<F, int(DIM)> g(f);
        b200::RegionsHandle regs;
        IntegratorAdaptiveIterations<Rule, EH>(eh, iterations).generate(ctx, g, range, regs);
        if constexpr (!std::is_same<Logger, LoggerNull>::value) logger.log(b200::download_regions<DIM>(ctx, regs.r));
        vb200_cv_params p; std::memset(&p, 0, sizeof(p));
        p.domain = b200::make_domain(range, res); p.shard = b200::current_shard(); p.spp = spp; p.seed = seed_;
        p.weight_strategy = weight_strategy; p.alpha = alpha; p.rr_policy = rr; p.rs_policy = rs; p.rs_power = rs_power; p.rs_cutoff = rs_cutoff;
        std::vector<float> flat(b200::bin_count(res), 0.0f);
        ctx.check(vb200_cv_integrate(ctx.get(), g.c_abi(), regs.r, &p, flat.data(), VB200_HOST, nullptr, nullptr));
        b200::apply_bins<false>(bins, res, flat);
        logger.log_progress(std::size_t(1), std::size_t(1));
    }
};
using IntegratorCrespo2021 = IntegratorAdaptiveVarianceReduction<Nested<Simpson,Trapezoidal>, error_heuristic_size<error_metric_relative>>;
inline IntegratorCrespo2021 integrator_crespo2021(std::size_t iterations, std::size_t spp, std::size_t seed = 0, std::size_t = 16) {
    return IntegratorCrespo2021(error_heuristic_size<error_metric_relative>(error_metric_relative(), 1.e-5), iterations, spp, seed);
}
// the reference's overloads that take a seed (integrator-adaptive-variance-reduction.h:21-49); RR = rr_uniform_region | rr_integral_region |
// rr_error_region | rr_pdf_region | region_stratification_uniform, CV = cv_optimize_weight | cv_fixed_weight(alpha), RS = region_sampling_uniform | _importance |
// _mis(power, cutoff) | _russian_roulette (the ...optimized factories of integrator-adaptive-variance-reduction-optimized.h:11-30 are the same call with
// region_stratification_uniform() in the RR slot)
template<typename RR, typename CV, typename RS, typename R, typename EH, typename = decltype(RR::id), typename = decltype(CV::id), typename = decltype(EH::id), typename = decltype(RS::id)>
auto integrator_adaptive_variance_reduction_parallel(const R&, const EH& eh, std::size_t iterations, const RR&, const CV& cv, const RS& rs, unsigned long spp, std::size_t seed = 0, std::size_t = 16) {
    return IntegratorAdaptiveVarianceReduction<R, EH>(eh, iterations, spp, seed, CV::id, cv.alpha, RR::id, RS::id, rs.power, rs.cutoff);
}
template<typename RR, typename CV, typename RS, typename R, typename EH, typename = decltype(RR::id), typename = decltype(CV::id), typename = decltype(EH::id), typename = decltype(RS::id)>
auto integrator_adaptive_variance_reduction_parallel_optimized(const R& r, const EH& eh, std::size_t iterations, const RR& rr, const CV& cv, const RS& rs,
        unsigned long spp, std::size_t seed = 0, std::size_t n = 16) {
    return integrator_adaptive_variance_reduction_parallel(r, eh, iterations, rr, cv, rs, spp, seed, n);
}
template<typename RR, typename CV, typename R, typename EH, typename = decltype(RR::id), typename = decltype(CV::id), typename = decltype(EH::id)>
auto integrator_adaptive_variance_reduction_parallel(const R&, const EH& eh, std::size_t iterations, const RR&, const CV& cv, unsigned long spp, std::size_t seed = 0, std::size_t = 16) {
    return IntegratorAdaptiveVarianceReduction<R, EH>(eh, iterations, spp, seed, CV::id, cv.alpha, RR::id);
}
template<typename RR, typename CV, typename RS, typename R, typename = decltype(RR::id), typename = decltype(CV::id), typename = decltype(R::id), typename = decltype(RS::id)>
auto integrator_adaptive_variance_reduction_parallel(const R&, std::size_t iterations, const RR&, const CV& cv, const RS& rs, unsigned long spp, std::size_t seed = 0, std::size_t = 16) {
    using EH = error_heuristic_default<error_metric_absolute>;
    return IntegratorAdaptiveVarianceReduction<R, EH>(EH(error_metric_absolute()), iterations, spp, seed, CV::id, cv.alpha, RR::id, RS::id, rs.power, rs.cutoff);
}
template<typename RR, typename CV, typename R, typename = decltype(RR::id), typename = decltype(CV::id), typename = decltype(R::id)>
auto integrator_adaptive_variance_reduction_parallel(const R&, std::size_t iterations, const RR&, const CV& cv, unsigned long spp, std::size_t seed = 0, std::size_t = 16) {
    using EH = error_heuristic_default<error_metric_absolute>;
    return IntegratorAdaptiveVarianceReduction<R, EH>(EH(error_metric_absolute()), iterations, spp, seed, CV::id, cv.alpha, RR::id);
}

// ---- Fubini family (reference src/combination/fubini.h:18-101, regions-generator-fubini.h:7-28, integrator-crespo2021.h:24-44) ---
template<std::size_t N, typename Float, std::size_t DIM>
std::tuple<Range<Float,N>, Range<Float,DIM-N>> range_split_at(const Range<Float,DIM>& range) {
    std::array<Float,N> a, b; std::array<Float,DIM-N> ra, rb;
    for (std::size_t i = 0; i < N; ++i) { a[i] = range.min(i); b[i] = range.max(i); }
    for (std::size_t i = N; i < DIM; ++i) { ra[i-N] = range.min(i); rb[i-N] = range.max(i); }
    return std::tuple<Range<Float,N>, Range<Float,DIM-N>>(Range<Float,N>(a,b), Range<Float,DIM-N>(ra,rb));
}
template<std::size_t N, typename Float>
std::tuple<Range<Float,N>, RangeInfinite<Float>> range_split_at(const RangeInfinite<Float>& range) {
    std::array<Float,N> a, b;
    for (std::size_t i = 0; i < N; ++i) { a[i] = range.min(i); b[i] = range.max(i); }
    std::vector<Float> ra = range.min(), rb = range.max();
    ra.erase(ra.begin(), ra.begin() + std::min(N, ra.size())); rb.erase(rb.begin(), rb.begin() + std::min(N, rb.size()));
    return std::tuple<Range<Float,N>, RangeInfinite<Float>>(Range<Float,N>(a,b), RangeInfinite<Float>(ra,rb));
}
namespace b200 {
// function_split_and_integrate_at<N>(f, monte_carlo(m, seed), range_rest) as a device functor (device/fubini.cuh)
template<std::size_t N, typename F, std::size_t R>
auto fubini_function(const F& f, const Range<float,R>& rest, unsigned long m, std::size_t seed) {
    return make_fubini_finite<int(N), int(R)>(f, rest.min().data(), rest.max().data(), m, seed);
}
template<std::size_t N, typename F>
auto fubini_function(const F& f, const RangeInfinite<float>& rest, unsigned long m, std::size_t seed) {
    std::vector<float> lo, hi; const std::size_t n = std::max(rest.min().size(), rest.max().size());
    if (n > VB200_MAX_DIM) throw std::runtime_error("viltrum_b200: at most 8 explicit entries in an infinite range");
    for (std::size_t i = 0; i < n; ++i) { lo.push_back(rest.min(i)); hi.push_back(rest.max(i)); }
    return make_fubini_infinite<int(N)>(f, lo.data(), hi.data(), int(n), m, seed);
}
}
// integrator_fubini<N>(first, monte_carlo(m, seed)) — fubini.h:78-101: `first` integrates g(x) = MC estimate of the rest integral
template<typename First, std::size_t N> class IntegratorFubini {
    First first; MonteCarlo rest;
public:
    IntegratorFubini(const First& f, const MonteCarlo& r) : first(f), rest(r) {}
    template<typename Bins, std::size_t DIMBINS, typename F, typename R, typename Logger>
    void integrate(Bins& bins, const std::array<std::size_t,DIMBINS>& res, const F& f, const R& range, Logger& logger) const {
        static_assert(N >= DIMBINS, "Fubini does not work with that many dimensions on bin resolution");      // fubini.h:88
        auto split = range_split_at<N>(range);
        first.integrate(bins, res, b200::fubini_function<N>(f, std::get<1>(split), rest.sample_count(), rest.seed()), std::get<0>(split), logger);
    }
};
template<std::size_t N, typename First> IntegratorFubini<First,N> integrator_fubini(const First& first, const MonteCarlo& rest) { return IntegratorFubini<First,N>(first, rest); }

// integrator_crespo2021_infinite<N>(iterations, mc_samples, spp, seed) — integrator-crespo2021.h:24-44 ('=')
template<std::size_t N> class IntegratorCrespo2021Infinite {
    std::size_t iterations, mc_samples, spp, seed_; int rr = VB200_RR_UNIFORM;
public:
    IntegratorCrespo2021Infinite(std::size_t it, std::size_t m, std::size_t s, std::size_t seed, int rr_policy = VB200_RR_UNIFORM) : iterations(it), mc_samples(m),
            spp(s), seed_(seed), rr(rr_policy) {}
    template<typename Bins, std::size_t DIMBINS, typename F, typename R, typename Logger>
    void integrate(Bins& bins, const std::array<std::size_t,DIMBINS>& res, const F& f, const R& range, Logger& logger) const {
        auto& ctx = b200::default_context();
        auto split = range_split_at<N>(range);
        const Range<float,N>& first = std::get<0>(split);
        // regions_generator_fubini<N>(adaptive heap, monte_carlo(mc_samples, 2*seed+1))
        auto g_gen = b200::fubini_function<N>(f, std::get<1>(split), mc_samples, 2*seed_+1);
        b200::Integrand<decltype(g_gen), int(N)> gen(g_gen);
        b200::RegionsHandle regs;
        using EH = error_heuristic_size<error_metric_relative>;
        IntegratorAdaptiveIterations<Nested<Simpson,Trapezoidal>, EH>(EH(error_metric_relative(), 1.e-5), iterations).generate(ctx, gen, first, regs);
        if constexpr (!std::is_same<Logger, LoggerNull>::value) logger.log(b200::download_regions<N>(ctx, regs.r));
        // residual: f through monte_carlo_per_bin(rng, 1) over the rest (regions-integrator-parallel-variance-reduction.h:69)
        auto g_res = b200::fubini_function<N>(f, std::get<1>(split), 1ul, seed_ + std::size_t(0x9E3779B97F4A7C15ull));
        b200::Integrand<decltype(g_res), int(N)> resid(g_res);
        vb200_cv_params p; std::memset(&p, 0, sizeof(p));
        p.domain = b200::make_domain(first, res); p.shard = b200::current_shard(); p.spp = spp; p.seed = seed_; p.rr_policy = rr;
        std::vector<float> flat(b200::bin_count(res), 0.0f);
        ctx.check(vb200_cv_integrate(ctx.get(), resid.c_abi(), regs.r, &p, flat.data(), VB200_HOST, nullptr, nullptr));
        b200::apply_bins<false>(bins, res, flat);
        logger.log_progress(std::size_t(1), std::size_t(1));
    }
};
template<std::size_t N> IntegratorCrespo2021Infinite<N> integrator_crespo2021_infinite(std::size_t iterations, std::size_t mc_samples, std::size_t spp, std::size_t seed = 0, std::size_t = 16) {
    return IntegratorCrespo2021Infinite<N>(iterations, mc_samples, spp, seed);
}

// integrator_adaptive_fubini_variance_reduction_parallel_optimized<N>(nested(simpson,trapezoidal), error_heuristic_size(relative), iterations, mc_samples,
// region_stratification_uniform(), cv_optimize_weight(), region_sampling_uniform(), spp, seed) — reference integrator-adaptive-fubini-variance-reduction-optimized.h:17-23
template<std::size_t N, typename R, typename EH, typename CV, typename RS>
IntegratorCrespo2021Infinite<N> integrator_adaptive_fubini_variance_reduction_parallel_optimized(const R&, const EH&, std::size_t iterations, unsigned long mc_samples,
        const region_stratification_uniform&,
                                                                                                  const CV&, const RS&, unsigned long spp, std::size_t seed = 0, std::size_t = 16) {
    static_assert(std::is_same<R, Nested<Simpson,Trapezoidal>>::value && std::is_same<EH, error_heuristic_size<error_metric_relative>>::value
            && CV::id == VB200_CV_OPTIMIZE_WEIGHT && RS::id == VB200_RS_UNIFORM,
                  "the device pipeline of the infinite-range control variates is built for the crespo2021 preset (simpson/trapezoidal, size/relative, cv_optimize_weight, region_sampling_uniform)");
    return IntegratorCrespo2021Infinite<N>(iterations, mc_samples, spp, seed, VB200_RR_STRATIFIED);
}

// ---- the front door (reference src/integrate.h:72-173) ------------------------------------------------------------------------
template<typename Integrator, typename Bins, std::size_t DIMBINS, typename F, typename Float, std::size_t DIM, typename Logger>
void integrate(const Integrator& integrator, Bins& bins, const std::array<std::size_t,DIMBINS>& resolution, const F& function, const Range<Float,DIM>& range, Logger& logger) {
    integrator.integrate(bins, resolution, b200::adapt<DIM>(function), range, logger);
}
template<typename Integrator, typename Bins, std::size_t DIMBINS, typename F, typename Float, typename Logger>
void integrate(const Integrator& integrator, Bins& bins, const std::array<std::size_t,DIMBINS>& resolution, const F& function, const RangeInfinite<Float>& range, Logger& logger) {
    integrator.integrate(bins, resolution, function, range, logger);
}
template<typename Integrator, typename Bins, std::size_t DIMBINS, typename F, typename Float, std::size_t DIM>
void integrate(const Integrator& integrator, Bins& bins, const std::array<std::size_t,DIMBINS>& resolution, const F& function, const Range<Float,DIM>& range) {
    LoggerNull log; integrate(integrator, bins, resolution, function, range, log);
}
template<typename Integrator, typename Bins, std::size_t DIMBINS, typename F, typename Float>
void integrate(const Integrator& integrator, Bins& bins, const std::array<std::size_t,DIMBINS>& resolution, const F& function, const RangeInfinite<Float>& range) {
    LoggerNull log; integrate(integrator, bins, resolution, function, range, log);
}
// single value (integrate.h:105-130): one bin
template<typename Integrator, typename F, typename Float, std::size_t DIM, typename Logger>
float integrate(const Integrator& integrator, const F& function, const Range<Float,DIM>& range, Logger& logger) {
    float sol(0.0f);
    b200::SingleBin bins{&sol};
    std::array<std::size_t,1> res{1};
    integrate(integrator, bins, res, function, range, logger);
    return sol;
}
template<typename Integrator, typename F, typename Float, typename Logger>
float integrate(const Integrator& integrator, const F& function, const RangeInfinite<Float>& range, Logger& logger) {
    float sol(0.0f);
    b200::SingleBin bins{&sol};
    std::array<std::size_t,1> res{1};
    integrate(integrator, bins, res, function, range, logger);
    return sol;
}
template<typename Integrator, typename F, typename R>
float integrate(const Integrator& integrator, const F& function, const R& range) { LoggerNull log; return integrate(integrator, function, range, log); }
// std::vector bins (integrate.h:132-137,169-173)
template<typename Integrator, typename T, typename F, typename R, typename Logger>
void integrate(const Integrator& integrator, std::vector<T>& bins, const F& function, const R& range, Logger& logger) {
    auto b = [&bins] (const std::array<std::size_t,1>& i) -> T& { return bins[i[0]]; };
    std::array<std::size_t,1> res{bins.size()};
    integrate(integrator, b, res, function, range, logger);
}
// nested std::vector bins (integrate.h:139-167; like upstream these two exist only with a logger): ragged rows are resized to the longest
template<typename Integrator, typename T, typename F, typename R, typename Logger>
void integrate(const Integrator& integrator, std::vector<std::vector<T>>& bins, const F& function, const R& range, Logger& logger) {
    auto b = [&bins] (const std::array<std::size_t,2>& i) -> T& { return bins[i[0]][i[1]]; };
    std::size_t max1 = 0;
    for (const std::vector<T>& v : bins) if (v.size() > max1) max1 = v.size();
    for (std::vector<T>& v : bins) v.resize(max1);
    std::array<std::size_t,2> res{bins.size(), max1};
    integrate(integrator, b, res, function, range, logger);
}
template<typename Integrator, typename T, typename F, typename R, typename Logger>
void integrate(const Integrator& integrator, std::vector<std::vector<std::vector<T>>>& bins, const F& function, const R& range, Logger& logger) {
    auto b = [&bins] (const std::array<std::size_t,3>& i) -> T& { return bins[i[0]][i[1]][i[2]]; };
    std::size_t max1 = 0, max2 = 0;
    for (const auto& v : bins) if (v.size() > max1) max1 = v.size();
    for (auto& v : bins) v.resize(max1);
    for (const auto& vv : bins) for (const auto& v : vv) if (v.size() > max2) max2 = v.size();
    for (auto& vv : bins) for (auto& v : vv) v.resize(max2);
    std::array<std::size_t,3> res{bins.size(), max1, max2};
    integrate(integrator, b, res, function, range, logger);
}
template<typename Integrator, typename T, typename F, typename R>
void integrate(const Integrator& integrator, std::vector<T>& bins, const F& function, const R& range) { LoggerNull log; integrate(integrator, bins, function, range, log); }

} // namespace viltrum
