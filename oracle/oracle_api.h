// TEST INFRASTRUCTURE — not product code.
//
// One C ABI, exported by BOTH CPU checkers so the tests can run them side by side:
//   oracle/_ref/libviltrum_ref.so  vo_kind() == "reference"  (unmodified reference headers, ref_harness_*.cpp)
//   oracle/liboracle.so            vo_kind() == "port"       (plain restatement, oracle.cpp)
// All bins / per-bin records are flat arrays in viltrum::tensor layout (dim-0-fastest,
// reference src/tensor.h:17-23).  Every function returns 0 on success, <0 on error
// (-1 unknown integrand, -2 unsupported dimension/rule combination, -3 buffer too small).
// Optional output pointers may be NULL.
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* vo_kind(void);
// dimension of a named integrand; -1 for infinite-dimensional ones, 0 if unknown
int vo_integrand_dim(const char* integrand);

// reference monte_carlo_per_bin_parallel(spp,seed) — src/monte-carlo/monte-carlo-per-bin-parallel.h:41-71
// bins: in/out, accumulated with '+='.  rec_samples [nbins*spp*dim], rec_sum/rec_sum2 [nbins] = sum f, sum f^2.
int vo_mc_per_bin_parallel(const char* integrand, int dimbins, const uint64_t* res,
                           const float* rmin, const float* rmax, uint64_t spp, uint64_t seed,
                           float* bins, float* rec_samples, double* rec_sum, double* rec_sum2);

// reference integrator_per_bin_parallel(monte_carlo(spp,seed)) — src/integrator-per-bin-parallel.h:16-35,
// src/monte-carlo/monte-carlo.h:39-63.  bins overwritten ('=').
int vo_per_bin_parallel_mc(const char* integrand, int dimbins, const uint64_t* res,
                           const float* rmin, const float* rmax, uint64_t spp, uint64_t seed,
                           float* bins, float* rec_samples, double* rec_sum, double* rec_sum2);

// reference monte_carlo(samples,seed) global scatter — src/monte-carlo/monte-carlo.h:39-63.  '+='.
int vo_monte_carlo(const char* integrand, int dimbins, const uint64_t* res,
                   const float* rmin, const float* rmax, uint64_t samples, uint64_t seed,
                   float* bins, float* rec_samples);

// reference monte_carlo_per_bin_parallel(spp,seed) over RangeInfinite —
// src/monte-carlo/monte-carlo-per-bin-parallel.h:73-100, src/monte-carlo/random-sequence-ref-dis.h:11-44.
// rmin/rmax have nrange entries (implicit [0,1] tail).  rec_len [nbins*spp] = sequence elements the
// integrand consumed for that path; rec_elems (capacity rec_cap floats) = those elements back to back in
// bin-major (tensor order), sample-minor order; *rec_used = floats needed.
int vo_mc_per_bin_parallel_inf(const char* integrand, int dimbins, const uint64_t* res,
                               const float* rmin, const float* rmax, int nrange,
                               uint64_t spp, uint64_t seed, float* bins,
                               double* rec_sum, double* rec_sum2,
                               uint32_t* rec_len, float* rec_elems, uint64_t rec_cap, uint64_t* rec_used);

// reference integrator_per_bin_parallel(monte_carlo(spp,seed)) over RangeInfinite — src/integrator-per-bin-parallel.h:16-35 +
// src/monte-carlo/monte-carlo.h:65-84 + src/monte-carlo/random-sequence-rng.h:11-45 (a fresh mt19937 per sample).  '='.
// Records as vo_mc_per_bin_parallel_inf (the elements f consumed through its own begin()).
int vo_per_bin_parallel_mc_inf(const char* integrand, int dimbins, const uint64_t* res,
                               const float* rmin, const float* rmax, int nrange,
                               uint64_t spp, uint64_t seed, float* bins,
                               double* rec_sum, double* rec_sum2,
                               uint32_t* rec_len, float* rec_elems, uint64_t rec_cap, uint64_t* rec_used);

// reference monte_carlo(samples,seed) over RangeInfinite, global scatter — src/monte-carlo/monte-carlo.h:65-84.  '+='.
int vo_monte_carlo_inf(const char* integrand, int dimbins, const uint64_t* res,
                       const float* rmin, const float* rmax, int nrange, uint64_t samples, uint64_t seed, float* bins);

// reference integrator_newton_cotes(rule) — src/newton-cotes/newton-cotes.h:11-14. rule in
// {"trapezoidal","simpson","boole"}.  '+='.
int vo_newton_cotes(const char* integrand, const char* rule, int dimbins, const uint64_t* res,
                    const float* rmin, const float* rmax, float* bins);

// reference integrator_adaptive_iterations(nested(H,L), heuristic, iterations) —
// src/nested/integrator-adaptive-iterations.h:12-15, src/nested/regions-generator-adaptive-heap.h:18-45,
// src/newton-cotes/regions-integrator-sequential.h:38-58.
//   rule      in {"simpson_trapezoidal","boole_simpson"}
//   heuristic in {"default_absolute","default_relative","size_absolute","size_relative"} or "mixed_<bins metric>_<rest metric>" =
//             error_heuristic_mixed (error-heuristic.h:49-98) whose remaining constructor arguments come from vo_set_mixed (reference defaults
//             until it is called); reg_err then holds the double key rounded to float
//   size_weight only used by size_* and mixed_* (reference defaults 1e-5 / 1e-3)
// Region list (iterations+1 regions, in the reference's heap-array order):
//   reg_min/reg_max [n*dim], reg_err [n], reg_dim [n], reg_data [n*S^dim] (dim-0-fastest samples).
int vo_adaptive_iterations(const char* integrand, const char* rule, const char* heuristic, double size_weight,
                           uint64_t iterations, int dimbins, const uint64_t* res,
                           const float* rmin, const float* rmax, float* bins,
                           float* reg_min, float* reg_max, float* reg_err, uint32_t* reg_dim, float* reg_data);

void vo_set_mixed(int dimension, double bins_weight, double size_threshold_bins, double size_threshold_rest, double error_increase_factor);

// reference integrator_adaptive_tolerance(nested(H,L), heuristic, tolerance) — src/nested/integrator-adaptive-tolerance.h:15-39: depth-first
// recursion, a region is integrated into the bins ('+=', sequential integrator) as soon as its heuristic error drops below the
// tolerance, otherwise split along the heuristic's dimension (child 0 first).  *nleaves = number of integrated regions.
// The port also returns the leaves in visiting order (reg_* with capacity reg_cap regions; may be NULL); the reference harness
// only counts them (the reference never hands its leaves to the logger) and leaves reg_* untouched.  Returns -3 if the leaf list
// exceeds reg_cap (bins and *nleaves are still complete), -4 if the recursion exceeds 128 levels.
int vo_adaptive_tolerance(const char* integrand, const char* rule, const char* heuristic, double size_weight, float tolerance,
                          int dimbins, const uint64_t* res, const float* rmin, const float* rmax, float* bins, uint64_t* nleaves,
                          uint64_t reg_cap, float* reg_min, float* reg_max, float* reg_err, uint32_t* reg_dim, float* reg_data);

// Double-precision twins (Range<double,DIM>, double integrand, double bins) of vo_newton_cotes / vo_adaptive_iterations.
// Integrands: "x2y2", "ind2", "cubic1", "poly3", "smooth_edge2", "shade4_16" (the double family of integrands.h).
int vo_newton_cotes_f64(const char* integrand, const char* rule, int dimbins, const uint64_t* res,
                        const double* rmin, const double* rmax, double* bins);
int vo_adaptive_iterations_f64(const char* integrand, const char* rule, const char* heuristic, double size_weight,
                               uint64_t iterations, int dimbins, const uint64_t* res,
                               const double* rmin, const double* rmax, double* bins,
                               double* reg_min, double* reg_max, double* reg_err, uint32_t* reg_dim, double* reg_data);

// reference integrator_crespo2021(iterations,spp,seed) — src/control-variates/integrator-crespo2021.h:7-22,
// src/control-variates/regions-integrator-parallel-variance-reduction.h:32-109.  bins overwritten ('=').
// Records (per bin, tensor order): rec_nregions [nbins]; rec_approx [nbins] (control-variate integral);
// rec_chosen [nbins*spp] index into the region list; rec_samples [nbins*spp*dim].
// Region list outputs as in vo_adaptive_iterations (rule = simpson_trapezoidal, S=3).
int vo_crespo2021(const char* integrand, uint64_t iterations, uint64_t spp, uint64_t seed,
                  int dimbins, const uint64_t* res, const float* rmin, const float* rmax, float* bins,
                  uint32_t* rec_nregions, float* rec_approx, uint32_t* rec_chosen, float* rec_samples,
                  float* reg_min, float* reg_max, float* reg_err, uint32_t* reg_dim, float* reg_data);

// The same integrator with cv_fixed_weight(alpha) instead of cv_optimize_weight (reference src/control-variates/weight-strategy.h:7-35):
// integrator_adaptive_variance_reduction_parallel(nested(simpson,trapezoidal), error_heuristic_size(relative,1e-5), iterations,
// rr_uniform_region(), cv_fixed_weight(alpha), region_sampling_uniform(), spp, seed).  Records as vo_crespo2021 (rec_approx: port only).
int vo_cv_fixed_weight(const char* integrand, uint64_t iterations, uint64_t spp, uint64_t seed, double alpha,
                       int dimbins, const uint64_t* res, const float* rmin, const float* rmax, float* bins,
                       uint32_t* rec_nregions, float* rec_approx, uint32_t* rec_chosen, float* rec_samples);

// The same integrator with the Russian-roulette policy and the weight strategy chosen by the caller:
// integrator_adaptive_variance_reduction_parallel(nested(simpson,trapezoidal), error_heuristic_size(relative,1e-5), iterations,
// RR, CV, region_sampling_uniform(), spp, seed) with RR = rr_uniform_region (0) / rr_integral_region (1) / rr_error_region (2) / rr_pdf_region (3)
// (reference src/control-variates/region-russian-roulette.h:9-147) and CV = cv_optimize_weight (weight_strategy 0) / cv_fixed_weight(alpha) (1).
int vo_cv_policies(const char* integrand, uint64_t iterations, uint64_t spp, uint64_t seed, int rr_policy, int weight_strategy, double alpha,
                   int dimbins, const uint64_t* res, const float* rmin, const float* rmax, float* bins,
                   uint32_t* rec_nregions, float* rec_approx, uint32_t* rec_chosen, float* rec_samples);

// ---- Fubini family ---------------------------------------------------------------------------------
// The integrand's first `nfirst` dimensions are handled by a "first" integrator over g(x) = integral of f(x, rest) estimated by
// monte_carlo(mc_samples, mc_seed) over the remaining dimensions — finite (named finite integrand, dim > nfirst, rmin/rmax have
// dim entries, nrange ignored) or infinite (named sequence integrand, rmin/rmax have nrange explicit entries).
// reference src/combination/fubini.h:51-101 (function_split_and_integrate_at, IntegratorFubini).
//   first = integrator_adaptive_iterations(nested rule, heuristic, iterations)        '+='
int vo_fubini_adaptive_mc(const char* integrand, int nfirst, const char* rule, const char* heuristic, double size_weight,
                          uint64_t iterations, uint64_t mc_samples, uint64_t mc_seed, int dimbins, const uint64_t* res,
                          const float* rmin, const float* rmax, int nrange, float* bins);
//   first = monte_carlo_per_bin_parallel(spp, seed)                                    '+='
int vo_fubini_mc_mc(const char* integrand, int nfirst, uint64_t spp, uint64_t seed, uint64_t mc_samples, uint64_t mc_seed,
                    int dimbins, const uint64_t* res, const float* rmin, const float* rmax, int nrange, float* bins);
// reference integrator_crespo2021_infinite<nfirst>(iterations, mc_samples, spp, seed) — src/control-variates/integrator-crespo2021.h:24-44,
// src/combination/regions-generator-fubini.h:7-28; the residual pass evaluates f through monte_carlo_per_bin(rng,1) over the
// rest (regions-integrator-parallel-variance-reduction.h:69).  '='.
int vo_crespo2021_infinite(const char* integrand, int nfirst, uint64_t iterations, uint64_t mc_samples, uint64_t spp, uint64_t seed,
                           int dimbins, const uint64_t* res, const float* rmin, const float* rmax, int nrange, float* bins);

// Multi-threaded CPU baseline used by bench.py (--impl reference / cpu_baseline): slabs the bin grid along the
// LAST bin dimension over nthreads std::threads, each calling the single-threaded entry point above on its
// slab with seed+slab.  path in {"mc_per_bin_parallel","per_bin_parallel_mc",
// "mc_per_bin_parallel_inf"}.
int vo_mt_per_bin(const char* path, const char* integrand, int dimbins, const uint64_t* res,
                  const float* rmin, const float* rmax, int nrange, uint64_t spp, uint64_t seed,
                  int nthreads, float* bins);


// ---- remaining control-variate policies (reference builds only: no port restatement — Simpson::sample inverts a cubic CDF with pow/acos/cos, so
// only statistical parity is on offer and the tests compare against the unmodified reference directly) ------------------------------------------
// integrator_region_based(regions_generator_adaptive_heap(nested(simpson,trapezoidal), size/relative 1e-5, iterations),
//   regions_integrator_parallel_variance_reduction(rr_uniform_region(), cv_optimize_weight(), RS, mt19937(seed), spp)) with RS = region_sampling_uniform (0) /
//   _importance (1) / _mis(power, cutoff) (2) / _russian_roulette (3) — reference src/control-variates/region-sampling.h:9-135.  '='.
int vo_cv_sampling(const char* integrand, uint64_t iterations, uint64_t spp, uint64_t seed, int rs_policy, double power, double cutoff,
                   int dimbins, const uint64_t* res, const float* rmin, const float* rmax, float* bins);
// integrator_adaptive_fubini_variance_reduction_parallel_optimized<nfirst>(nested(simpson,trapezoidal), size/relative 1e-5, iterations, mc_samples,
//   region_stratification_uniform(), cv_optimize_weight(), region_sampling_uniform(), spp, seed) over a sequence integrand and an infinite range —
//   reference src/control-variates/integrator-adaptive-fubini-variance-reduction-optimized.h:17-23, regions-integrator-parallel-variance-reduction-optimized.h:70-142.  '='.
int vo_cv_optimized_infinite(const char* integrand, int nfirst, uint64_t iterations, uint64_t mc_samples, uint64_t spp, uint64_t seed,
                             int dimbins, const uint64_t* res, const float* rmin, const float* rmax, int nrange, float* bins);

// ---- timing legs (reference builds only; the port does not export them) -----------------------------------------------------
// vo_set_threads: threads behind the reference's std::for_each(par_unseq, ...) loops in oracle/_ref/libviltrum_ref_mt.so (the same
// harness over the same unmodified reference, with oracle/pstl_threads/execution as the parallel-STL back end upstream takes from
// TBB); returns the value in effect (always 1 in the serial build).  vo_phase_times: t[0] = seconds until the last region-based call
// handed its region list to Logger::log (generation), t[1] = seconds of the whole call.
int  vo_set_threads(int n);
void vo_phase_times(double* t);

#ifdef __cplusplus
}
#endif
