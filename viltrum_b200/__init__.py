"""viltrum_b200 — B200-native per-bin integration hot path of adolfomunoz/viltrum behind the reference's own
vocabulary.  The product is the CUDA library (viltrum_b200/csrc -> libviltrum_b200.so, C ABI in
include/viltrum_b200.h) and the C++17 drop-in headers (include/viltrum_b200/viltrum.h); this Python package is
the thin host-side mirror used by the tests and the benchmark (ctypes over the C ABI, numpy/torch buffers)."""
from .host import (Context, Regions, integrate, monte_carlo, monte_carlo_per_bin_parallel, integrator_per_bin_parallel,  # noqa: F401
                   integrator_newton_cotes, integrator_adaptive_iterations, integrator_crespo2021, nested,
                   integrator_fubini, integrator_crespo2021_infinite, integrator_adaptive_tolerance, cv_fixed_weight, cv_optimize_weight, rr_uniform_region, rr_integral_region, rr_error_region, rr_pdf_region,
                   integrator_adaptive_variance_reduction_parallel, steps, FubiniIntegrand, range_split_at,
                   error_heuristic_default, error_heuristic_size, error_heuristic_mixed, error_metric_absolute, error_metric_relative,
                   range_primary, range_primary_infinite, Range, RangeInfinite, builtin_names, shard_for_rank, sample_shard_for_rank,
                   region_sampling_uniform, region_sampling_importance, region_sampling_mis, region_sampling_russian_roulette, region_stratification_uniform,
                   integrator_adaptive_fubini_variance_reduction_parallel_optimized, IntegratorAdaptiveIterations, RegionSampling)
from ._capi import Vb200Error  # noqa: F401
