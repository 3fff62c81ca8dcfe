/* viltrum_b200 — C ABI of the B200-native per-bin integration hot path (libviltrum_b200.so).
 *
 * This is the drop-in boundary SURVEY.md §8(b) describes: plain pointers, sizes and POD structs, no C++
 * types, no torch types.  The reference (adolfomunoz/viltrum, header-only C++17) has no FFI of its own; the
 * "binding" is the set of integrator classes whose
 *     integrate(bins, bin_resolution, f, range, logger) const
 * member viltrum::integrate dispatches to (reference src/integrate.h:72-90).  Each driver below replaces one
 * such member; include/viltrum_b200/viltrum.h re-creates the reference's C++ vocabulary on top of it
 * (see INTEGRATION.md).
 *
 * Conventions
 *   - one vb200_ctx per process and GPU (one process per GPU; multi-GPU = bin-grid sharding through
 *     vb200_shard, no data-path collective — SURVEY.md §8(e));
 *   - bins are flat float arrays in the reference's tensor layout, dimension 0 fastest
 *     (reference src/tensor.h:17-23), addressed from the base of the FULL grid even when a call only
 *     integrates a shard of it;
 *   - every pointer argument carries a memory-space flag: VB200_HOST pointers are staged through device
 *     scratch inside the call (the reference-facing, end-to-end path), VB200_DEVICE pointers are used in place
 *     (the resident path);
 *   - all entry points return VB200_OK (0) or a negative vb200_status; vb200_last_error() gives the text.
 *     The reference's hot path returns void and throws nothing (SURVEY.md §8b "Errors"); the C++ wrapper turns
 *     a non-zero status into std::runtime_error;
 *   - there is NO CPU fallback: without a CUDA device vb200_create fails with VB200_ERR_NO_DEVICE.
 *
 * Integrands are C++ functors compiled by nvcc in the USER's translation unit.  They cross this ABI as a
 * vb200_integrand: the functor's bytes (trivially copyable) plus a table of launch thunks instantiated from
 * the kernel templates in include/viltrum_b200/device/ (viltrum::b200::make_integrand<F,DIM>()).  The
 * library also ships the synthetic integrands of SURVEY.md §8(d) (vb200_builtin_integrand) so that C, Python
 * (ctypes) and the benchmark can drive every path without an nvcc TU of their own.
 */
#ifndef VILTRUM_B200_H
#define VILTRUM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB200_ABI_VERSION 3u
#define VB200_MAX_DIM      8     /* finite integrands: 1..8 dimensions (reference VILTRUM_MAX_DIMENSIONS_REGION = 6, region.h:16-18) */
#define VB200_MAX_DIMBINS  3     /* reference binned overloads go up to 3-D containers (integrate.h:132-167) */

typedef enum vb200_status {
    VB200_OK = 0,
    VB200_ERR_NO_DEVICE = -1,      /* no CUDA device / driver: the product never falls back to the CPU */
    VB200_ERR_INVALID = -2,        /* bad argument (dimension mismatch is a compile error upstream, integrate.h:75-77) */
    VB200_ERR_CUDA = -3,           /* a CUDA runtime call or kernel failed; see vb200_last_error */
    VB200_ERR_UNSUPPORTED = -4,    /* combination not implemented for this integrand (missing thunk) */
    VB200_ERR_NOMEM = -5
} vb200_status;

typedef enum vb200_mem { VB200_HOST = 0, VB200_DEVICE = 1 } vb200_mem;

typedef struct vb200_ctx vb200_ctx;

/* ---- context ------------------------------------------------------------------------------------------ */
int         vb200_create(int device, vb200_ctx** out);
void        vb200_destroy(vb200_ctx* ctx);
const char* vb200_last_error(const vb200_ctx* ctx);      /* ctx may be NULL: error of the last failed vb200_create */
/* stream all work of this context is enqueued on (a cudaStream_t); calls with VB200_HOST outputs synchronise it
 * before returning, calls whose outputs are all VB200_DEVICE return with the work enqueued. */
void*       vb200_stream(vb200_ctx* ctx);
int         vb200_synchronize(vb200_ctx* ctx);
int         vb200_sm_count(const vb200_ctx* ctx);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
uint64_t    vb200_launch_count(const vb200_ctx* ctx);
/* Per-kernel device time of the residual-sampling kernel of vb200_cv_integrate (the heaviest kernel of the control-variate pipeline; its
 * roofline in bench.py needs the kernel's own duration, not the step's).  While enabled, every launch of that kernel is bracketed by a pair
 * of CUDA events on the context's stream (no synchronisation, two event records per launch); vb200_kernel_timer_read synchronises the
 * stream, returns the summed duration and the number of launches since the last read, and releases the events. */
int         vb200_kernel_timer(vb200_ctx* ctx, int enable);
int         vb200_kernel_timer_read(vb200_ctx* ctx, double* ms_total, uint64_t* launches);

/* Pin and map a caller-owned host buffer (cudaHostRegister, mapped + portable) for the lifetime of the registration.  The sampling
 * drivers (vb200_mc_per_bin, vb200_mc_per_bin_inf) recognise VB200_HOST bins that lie inside a registered buffer and let the kernel
 * apply the reference's '+=' / '=' to them directly over PCIe (zero-copy read-modify-write): no staging buffer and no host-side pass,
 * which is what limits the end-to-end rate when several ranks share one host.  The buffer must stay allocated until
 * vb200_host_unregister / vb200_destroy.  Unregistered VB200_HOST bins keep working through the staged path. */
int         vb200_host_register(vb200_ctx* ctx, void* ptr, size_t bytes);
int         vb200_host_unregister(vb200_ctx* ctx, void* ptr);

/* Measured FP32 (non-tensor) peak of the device: a dependent-FFMA chain kernel (8 independent chains per thread, all SMs fully
 * occupied), timed with CUDA events on the context's stream; best of `reps`.  This is the roofline denominator bench.py reports
 * next to the nominal 2*128*SMs*clock figure (MEASURED_PEAKS.json carries HBM and bf16-tensor peaks only). */
int         vb200_measure_fp32_peak(vb200_ctx* ctx, int reps, double* tflops);

/* Host-side evaluation of the library's generators (include/viltrum_b200/device/philox.cuh, xoshiro.cuh, threefry.cuh), for
 * known-answer tests and for callers that want to predict which sample a (seed, bin, sample) triple maps to.
 * vb200_xoshiro128pp: advances `state` n times and stores the n outputs.  vb200_threefry4x32: the experiment generator of
 * profiles/exp/k1_mix.cu (rounds = 12, 13 or 20), not used by the shipped kernels. */
void        vb200_philox4x32_10(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]);
void        vb200_xoshiro128pp(uint32_t state[4], uint64_t n, uint32_t* out);
int         vb200_threefry4x32(int rounds, const uint32_t counter[4], const uint32_t key[4], uint32_t out[4]);

/* ---- integrands --------------------------------------------------------------------------------------- */
struct vb200_integrand;
/* launch thunk: enqueue one kernel instantiated for the integrand's functor type.  `args` points at the
 * launch struct of that kernel kind (below).  Returns a cudaError_t value (0 = success). */
typedef int (*vb200_launch_fn)(const struct vb200_integrand* self, const void* args, void* stream);

enum {
    VB200_K_MC_PER_BIN = 0,        /* vb200_mc_launch      : Philox per-bin sampler (fused draw/eval/reduce/store) */
    VB200_K_MC_REPLAY = 1,         /* vb200_replay_launch  : recorded samples, sequential reference arithmetic    */
    VB200_K_WALK = 2,              /* vb200_walk_launch    : infinite-dimensional lazy-sequence paths             */
    VB200_K_WALK_REPLAY = 3,       /* vb200_walk_replay_launch                                                     */
    VB200_K_EVAL_POINTS = 4,       /* vb200_eval_launch    : values[i] = f(points[i]) — region fill / batched splits / CV residual */
    VB200_K_ADAPTIVE_EXACT = 5,    /* vb200_greedy_launch  : persistent single-CTA greedy heap refinement (batch size 1) */
    VB200_K_MC_SCATTER = 6,        /* vb200_scatter_launch : global sampler scattering into bins (monte-carlo.h:39-63) */
    VB200_K_WALK_SCATTER = 7,      /* vb200_scatter_launch : the same over an infinite range (monte-carlo.h:65-84) */
    VB200_K_COUNT = 8
};

#define VB200_INTEGRAND_EXACT 1u   /* thunks were compiled with --fmad=false: bit-exact twin of a CPU build with -ffp-contract=off */
#define VB200_INTEGRAND_F64   2u   /* functor is double f(std::array<double,dim>): points/values of the eval thunk are doubles */

typedef struct vb200_integrand {
    uint32_t        abi_version;    /* VB200_ABI_VERSION */
    int32_t         dim;            /* >0: f(std::array<float,dim>) ; -1: f(sequence) over an infinite range */
    const void*     functor;        /* trivially copyable functor object, passed to kernels by value */
    uint32_t        functor_bytes;
    uint32_t        flags;          /* VB200_INTEGRAND_* */
    const char*     name;           /* for error messages */
    vb200_launch_fn launch[VB200_K_COUNT];   /* NULL = kind not available for this integrand */
} vb200_integrand;

/* Synthetic integrands compiled into the library: "x2y2", "ind2", "cubic1", "poly3",
 * "shade4_16", "shade4_64", "shade5_16", "shade5_64", "smooth_edge2", "walk", "decay".  exact != 0 selects the
 * --fmad=false instantiation.  Returns NULL if unknown. */
const vb200_integrand* vb200_builtin_integrand(const char* name, int exact);
int                    vb200_builtin_count(void);
const char*            vb200_builtin_name(int index);

/* Fubini adapter over a built-in integrand (reference src/combination/fubini.h:51-75,
 * function_split_and_integrate_at<nfirst>(f, monte_carlo(mc_samples, seed), range_rest)): the returned descriptor is an
 * nfirst-dimensional integrand  g(x) = vol(rest)/mc_samples * sum_s f(x (+) r_s)  with r_s uniform in the rest range —
 * rest_min/rest_max have nrest entries: exactly dim-nfirst for finite integrands, 0..VB200_MAX_DIM explicit entries (implicit [0,1]
 * beyond) for sequence integrands.  It goes wherever an integrand goes: vb200_mc_per_bin (integrator_fubini<N>(monte_carlo_per_bin, ..)),
 * vb200_regions_generate_adaptive (regions_generator_fubini<N>), vb200_cv_integrate with mc_samples = 1 (the residual pass of
 * integrator_crespo2021_infinite<N>, integrator-crespo2021.h:24-44).  Available: ("poly3",1|2) ("shade4_16"|"shade4_64",2)
 * ("shade5_16",2|3) ("decay",1|2) ("walk",1|2).  NULL if unknown.  Free with vb200_integrand_free (a no-op for every other descriptor).
 * C++ callers wrap their own functors with viltrum::integrator_fubini / integrator_crespo2021_infinite (include/viltrum_b200/viltrum.h). */
const vb200_integrand* vb200_builtin_fubini(const char* name, int nfirst, const float* rest_min, const float* rest_max, int nrest,
                                            uint64_t mc_samples, uint64_t seed);
void                   vb200_integrand_free(const vb200_integrand* f);

/* ---- multi-GPU: one process per GPU, an optional NCCL communicator per context --------------------------- */
/* The per-bin paths shard over independent bins (vb200_shard) and need no communicator.  The two exchanges of the design — the
 * split-sample allreduce of vb200_monte_carlo (VB200_MC_ALLREDUCE) and vb200_regions_broadcast — run over a communicator the context
 * owns.  vb200_comm_unique_id (one rank) makes the 128-byte rendezvous token (ncclGetUniqueId); hand it to every rank by whatever
 * the host application has (torch.distributed / MPI broadcast, a file), then every rank calls vb200_comm_init(ctx, id, rank, world)
 * (ncclCommInitRank on the context's device; collective).  NCCL is bound at run time with dlopen("libnccl.so.2") — the copy the process
 * already carries (PyTorch's) or the system's; VB200_NCCL_LIB overrides the path; VB200_ERR_UNSUPPORTED if none is found. */
#define VB200_COMM_ID_BYTES 128
int vb200_comm_unique_id(vb200_ctx* ctx, void* id /* VB200_COMM_ID_BYTES out */);
int vb200_comm_init(vb200_ctx* ctx, const void* id, int rank, int world);
int vb200_comm_destroy(vb200_ctx* ctx);
int vb200_comm_rank(const vb200_ctx* ctx);
int vb200_comm_size(const vb200_ctx* ctx);      /* 1 without a communicator */
int vb200_nccl_version(void);                   /* ncclGetVersion of the bound library, 0 if NCCL could not be loaded */

/* ---- shared parameter blocks -------------------------------------------------------------------------- */
/* Integration box + bin grid.  For infinite ranges `dim` is the number of explicit entries of rmin/rmax
 * (reference RangeInfinite: implicit [0,1] tail, range-infinite.h:31-37) and may be 0. */
typedef struct vb200_domain {
    int32_t  dim;
    int32_t  dimbins;                      /* bins span the FIRST dimbins dimensions (README.md:69-71) */
    float    rmin[VB200_MAX_DIM];
    float    rmax[VB200_MAX_DIM];
    uint64_t res[VB200_MAX_DIMBINS];       /* bins per dimension */
    float    drange[VB200_MAX_DIMBINS];    /* filled by the library: (max-min)/float(res), the reference's bin pitch
                                              (monte-carlo-per-bin-parallel.h:46-47); callers may leave it 0 */
    int32_t  reserved;
} vb200_domain;

/* double-precision twin (Range<double,DIM>) for the Newton-Cotes region family */
typedef struct vb200_domain_f64 {
    int32_t  dim;
    int32_t  dimbins;
    double   rmin[VB200_MAX_DIM];
    double   rmax[VB200_MAX_DIM];
    uint64_t res[VB200_MAX_DIMBINS];
} vb200_domain_f64;

/* Bin-grid shard handled by one call/GPU: linear bin indices [begin,end) in tensor order.  {0,0} = whole grid (a zero-initialised
 * parameter block integrates everything).  begin == end != 0 is an EMPTY shard: the call does nothing — a rank that owns no rows, or an
 * empty sample share in vb200_monte_carlo; write an empty shard that starts at the origin as VB200_SHARD_EMPTY ({UINT64_MAX, UINT64_MAX}).
 * Random streams are keyed by the GLOBAL bin index, so results do not depend on how the grid is sharded. */
typedef struct vb200_shard { uint64_t begin, end; } vb200_shard;
#define VB200_SHARD_EMPTY_INDEX 0xffffffffffffffffull

typedef enum vb200_mc_flavor {
    VB200_MC_PER_BIN = 0,     /* monte_carlo_per_bin_parallel(spp,seed): bins(p) += sum f * vol(range)/spp
                                 (reference src/monte-carlo/monte-carlo-per-bin-parallel.h:41-71 and :73-100) */
    VB200_PER_BIN_MC = 1      /* integrator_per_bin_parallel(monte_carlo(spp,seed)): bins(p) = nbins * (sum f * vol(bin box)/spp)
                                 (reference src/integrator-per-bin-parallel.h:16-35 + src/monte-carlo/monte-carlo.h:39-63) */
} vb200_mc_flavor;

/* ---- per-bin Monte Carlo (SURVEY.md §8a rows a2,a3,a4) ------------------------------------------------- */
/* Sampler options of vb200_mc_per_bin (bit flags; 0 = the fastest configuration).
 *   Random streams.  Default: every (bin, lane sub-stream) owns an xoshiro128++ stream whose 128-bit state is
 *   Philox4x32-10(key = seed, counter = (bin lo, bin hi, sub-stream, 'strm')) — the reference's own scheme (a per-bin generator seeded
 *   from a master stream, monte-carlo-per-bin-parallel.h:50-58; xoshiro128++ is one of the generators it vendors, src/rng/XoshiroCpp.hpp:531)
 *   with a counter-based master, so a bin's samples do not depend on how the grid is sharded.  Add / rotate / xor only: the generator
 *   runs on the ALU pipe next to the integrand's FFMA2 (60 % of FP32 peak on C2 against 53 % with pure Philox, profiles/k1_rng_r2.txt).
 *   VB200_MC_RNG_PHILOX: every word is Philox4x32-10(key = seed, counter = (bin lo, bin hi, sample group, call)) — stateless, a sample's
 *   coordinates depend on (seed, bin, sample index) only.
 *   Sample lattice.  Default: coordinates of the binned dimensions carry 16 random bits inside the bin when every binned dimension of the
 *   WHOLE grid has >= 256 bins (the lattice along such a dimension still has >= 2^24 points over the range, the reference's
 *   generate_canonical<float,24> resolution), 24 bits otherwise; non-binned dimensions always 24.  VB200_MC_LATTICE24: 24 bits everywhere. */
#define VB200_MC_RNG_PHILOX 1
#define VB200_MC_LATTICE24  2
/* vb200_monte_carlo only — split-sample mode across the context's communicator (vb200_comm_init): this rank draws samples
 * [spp*rank/world, spp*(rank+1)/world) of the global sample counter, the partial grids are summed with ncclAllReduce over NVLink, and every
 * rank applies the '+=' of the TOTAL to its bins: the one allreduce of the design (SURVEY.md §8e "few bins, many samples").  shard stays {0,0}. */
#define VB200_MC_ALLREDUCE  4
enum { VB200_RNG_XOSHIRO = 0, VB200_RNG_PHILOX = 1 };      /* vb200_mc_launch.rng */

typedef struct vb200_mc_params {
    vb200_domain domain;
    vb200_shard  shard;
    uint64_t     spp;
    uint64_t     seed;          /* generator key (Philox key = (seed lo, seed hi)) */
    int32_t      flavor;        /* vb200_mc_flavor: also fixes the write semantics ('+=' vs '=') */
    int32_t      options;       /* VB200_MC_* flags above (vb200_mc_per_bin only; the other drivers ignore them) */
} vb200_mc_params;

/* bins: base of the full grid.  sum_f / sum_f2 (optional, may be NULL; same memory space as bins): raw per-bin
 * sum f and sum f^2 of the shard's bins, for the tests' 3-sigma gate. */
int vb200_mc_per_bin(vb200_ctx* ctx, const vb200_integrand* f, const vb200_mc_params* p,
                     float* bins, int bins_mem, float* sum_f, float* sum_f2);

/* Sample-replay mode (BASELINE.json north_star: "match exactly in a sample-replay mode that feeds the
 * reference's sample points").  samples: [nbins_of_shard][spp][dim] floats, bin-major in tensor order starting
 * at shard.begin.  One thread per bin accumulates sequentially with the reference's promotions
 * (float(double(acc)+double(f)*factor)); with an EXACT integrand the result is bit-identical to the reference. */
int vb200_mc_per_bin_replay(vb200_ctx* ctx, const vb200_integrand* f, const vb200_mc_params* p,
                            const float* samples, int samples_mem, float* bins, int bins_mem);

/* ---- infinite-dimensional paths (rows a5, a20) --------------------------------------------------------- */
/* domain.dim = explicit range entries (0..VB200_MAX_DIM); both flavors (VB200_PER_BIN_MC = row a20:
 * integrator_per_bin_parallel(monte_carlo) over RangeInfinite, monte-carlo.h:65-84, bins(p) = nbins * sum f * vol(bin box)/spp).
 * Sequence element i of sample s in bin b = u * (max_i - min_i) + min_i with u drawn from
 * Philox(key=seed, counter=(b, s, i/4))[i%4]; the first dimbins elements are confined to the bin. */
int vb200_mc_per_bin_inf(vb200_ctx* ctx, const vb200_integrand* f, const vb200_mc_params* p,
                         float* bins, int bins_mem, float* sum_f, float* sum_f2);

/* Replay of recorded lazy sequences: elems holds every path's elements back to back (bin-major, sample-minor),
 * offsets [nbins_of_shard*spp + 1] are prefix sums of the per-path lengths.  Reading past a path's recorded
 * length yields +inf (so Russian-roulette loops end) and sets a sticky error (VB200_ERR_INVALID on return). */
int vb200_mc_per_bin_inf_replay(vb200_ctx* ctx, const vb200_integrand* f, const vb200_mc_params* p,
                                const uint64_t* offsets, const float* elems, int mem, float* bins, int bins_mem);

/* ---- global Monte Carlo with scatter binning (row a6; SURVEY.md §8f "next" #1) -------------------------- */
/* monte_carlo(samples,seed): bins(pos(x)) += f(x) * nbins*vol/samples   (reference src/monte-carlo/monte-carlo.h:39-63).
 * Sequence integrands (infinite ranges, monte-carlo.h:65-84) take the bin from the first dimbins sequence elements.
 * shard selects a SAMPLE index range [begin,end) here (split-bin mode: every GPU draws part of the samples and
 * the caller sums the partial grids — the one allreduce the design allows). */
int vb200_monte_carlo(vb200_ctx* ctx, const vb200_integrand* f, const vb200_mc_params* p /* spp = total samples */,
                      float* bins, int bins_mem);

/* ---- region tables (rows a7-a14) ----------------------------------------------------------------------- */
typedef enum vb200_rule {
    VB200_RULE_TRAPEZOIDAL = 2, VB200_RULE_SIMPSON = 3, VB200_RULE_BOOLE = 5,       /* value = samples per dimension (rules.h) */
    VB200_RULE_SIMPSON_TRAPEZOIDAL = 32, VB200_RULE_BOOLE_SIMPSON = 53              /* nested(high,low) pairs (nested.h:7-34) */
} vb200_rule;
/* Steps<Q,N> composite rule (reference src/newton-cotes/rules.h:321-388): n pieces of rule q (q = 2, 3 or 5 samples) per dimension,
 * (q-1)*n+1 samples per dimension.  Fixed-rule integration only (vb200_regions_generate_single + vb200_regions_integrate_bins). */
#define VB200_RULE_STEPS(q, n) (0x1000000 | ((q) << 16) | ((n) & 0xffff))
#define VB200_RULE_IS_STEPS(rule) (((rule) & 0x1000000) != 0)
typedef enum vb200_heuristic { VB200_HEURISTIC_DEFAULT = 0, VB200_HEURISTIC_SIZE = 1,                     /* error-heuristic.h:10-46 */
                               VB200_HEURISTIC_MIXED = 2 } vb200_heuristic;                                 /* error_heuristic_mixed, error-heuristic.h:49-98 */
/* error_heuristic_mixed(metric_bins, metric_rest, dimension, bins_weight, size_weight, size_threshold_bins, size_threshold_rest, error_increase_factor):
 * vb200_adaptive_params.metric is the bins metric and .size_weight the size weight; the rest comes from this block (reference defaults:
 * dimension 2, bins_weight 1.0, size_weight 1e-3, thresholds 1/1024 and 1/16, increase factor 1e4).  Its keys are doubles upstream; the exact
 * mode (batch = 1) keeps them as doubles, the batched mode orders by the key rounded to float. */
typedef struct vb200_mixed_heuristic {
    int32_t metric_rest;            /* vb200_metric of the non-binned dimensions */
    int32_t dimension;              /* first dimension that takes the rest metric */
    double  bins_weight, size_threshold_bins, size_threshold_rest, error_increase_factor;
} vb200_mixed_heuristic;
typedef enum vb200_metric { VB200_METRIC_ABSOLUTE = 0, VB200_METRIC_RELATIVE = 1 } vb200_metric;         /* error-metric.h:10-41 */

/* Leaf table produced by a generator ("region tree" of north_star = this flat table).
 * Device resident, SoA; order is the reference's heap-array order in exact mode.
 * Lifetime: a table belongs to the context that made it.  Free it with vb200_regions_free before or after vb200_destroy — destroying
 * the context releases the device memory of every outstanding table and leaves the handles valid only for vb200_regions_free
 * (count/dim/samples then read 0 regions); every other use of a table after its context is gone is an error. */
typedef struct vb200_regions vb200_regions;

typedef struct vb200_adaptive_params {
    vb200_domain domain;        /* dimbins/res unused by the generator */
    int32_t  rule;              /* nested pair */
    int32_t  heuristic;         /* vb200_heuristic */
    int32_t  metric;            /* vb200_metric */
    int32_t  batch;             /* 1 = exact greedy order (reference regions-generator-adaptive-heap.h:18-45, bit-exact with an
                                   EXACT integrand); 0 = batched top-k refinement (throughput mode, auto batch size);
                                   >1 = batched with at most this many splits per round */
    double   size_weight;       /* error_heuristic_size weight (reference default 1e-5) */
    uint64_t iterations;        /* number of splits: the table ends with iterations+1 regions */
    vb200_mixed_heuristic mixed;/* VB200_HEURISTIC_MIXED only */
} vb200_adaptive_params;
/* Range<double,DIM>: the same generator over a VB200_INTEGRAND_F64 integrand — every sample, error and heap key a double, as upstream for
 * Float = double.  Exact greedy mode only (batch must be 1); the table is a double table (vb200_regions_*_f64). */
typedef struct vb200_adaptive_params_f64 {
    vb200_domain_f64 domain;
    int32_t  rule, heuristic, metric, batch;
    double   size_weight;
    uint64_t iterations;
    vb200_mixed_heuristic mixed;
} vb200_adaptive_params_f64;

int vb200_regions_generate_adaptive(vb200_ctx* ctx, const vb200_integrand* f, const vb200_adaptive_params* p, vb200_regions** out);
/* integrator_adaptive_tolerance(nested(h,l), heuristic, tolerance) — reference src/nested/integrator-adaptive-tolerance.h:15-39: regions
 * are split (along the heuristic's dimension) until their heuristic error is below `tolerance`.  The reference recurses depth first;
 * here every round splits all failing regions at once and the leaves are finally sorted into the reference's depth-first order by
 * their root-to-leaf path, so vb200_regions_integrate_bins on the result reproduces the reference's bins bit for bit (EXACT
 * integrand).  Fails with VB200_ERR_UNSUPPORTED past 128 levels of subdivision and VB200_ERR_NOMEM past max_regions. */
typedef struct vb200_tolerance_params {
    vb200_domain domain;        /* dimbins/res unused by the generator */
    int32_t  rule;              /* nested pair */
    int32_t  heuristic;         /* vb200_heuristic */
    int32_t  metric;            /* vb200_metric */
    float    tolerance;         /* the reference stores it as float (integrator-adaptive-tolerance.h:13) */
    double   size_weight;
    uint64_t max_regions;       /* 0 = 2^27 */
} vb200_tolerance_params;
int vb200_regions_generate_tolerance(vb200_ctx* ctx, const vb200_integrand* f, const vb200_tolerance_params* p, vb200_regions** out);
/* regions_generator_single (reference src/newton-cotes/regions-generator-single.h:12-20): one region over the range */
int vb200_regions_generate_single(vb200_ctx* ctx, const vb200_integrand* f, const vb200_domain* domain, int rule, vb200_regions** out);
/* upload an externally produced table (tests: the reference's own region list) */
int vb200_regions_upload(vb200_ctx* ctx, int dim, int rule, uint64_t count,
                         const float* rmin, const float* rmax, const float* err, const uint32_t* errdim, const float* data,
                         vb200_regions** out);
uint64_t vb200_regions_count(const vb200_regions* r);
int      vb200_regions_dim(const vb200_regions* r);
int      vb200_regions_samples(const vb200_regions* r);     /* S^dim values per region */
/* host copies, AoS like the reference's Logger::log view: rmin/rmax [n*dim], err [n], errdim [n], data [n*S^dim]; any may be NULL */
int      vb200_regions_download(vb200_ctx* ctx, const vb200_regions* r, float* rmin, float* rmax, float* err, uint32_t* errdim, float* data);
void     vb200_regions_free(vb200_regions* r);

/* RegionsIntegratorSequential (reference src/newton-cotes/regions-integrator-sequential.h:38-58):
 * bins(pos) += nbins * region.integral_subrange(bin ∩ region) for every region and touched bin
 * (pixels_in_region incl. its 0.99f rule, region.h:454-463).  Bin-major, regions visited in table order per bin,
 * so the float summation order — and the bits — match the reference.  Also serves RegionsIntegratorParallelRegions. */
int vb200_regions_integrate_bins(vb200_ctx* ctx, const vb200_regions* r, const vb200_domain* domain, const vb200_shard* shard,
                                 float* bins, int bins_mem);

/* Region-table broadcast over the context's communicator (SURVEY.md §8e "CV residual": one table, bins slabbed over the GPUs; reference
 * regions-integrator-parallel-variance-reduction.h:53-63 builds every bin's list from the one table): the root passes its table, every
 * other rank passes *r == NULL and receives a new single-precision table of the same shape (free it with vb200_regions_free).
 * Collective; enqueued on the context's stream.  Generating the (deterministic) table on every rank instead costs no exchange at all
 * and is what bench.py times by default — DESIGN.md §6 has both numbers. */
int vb200_regions_broadcast(vb200_ctx* ctx, vb200_regions** r, int root);

/* ---- double precision (north_star: Newton-Cotes within 1e-12 in fp64) ---------------------------------------------------- */
/* The same region family for Range<double,DIM>: integrands flagged VB200_INTEGRAND_F64 (functor over std::array<double,DIM>
 * returning double), double bins, every rule/fold/accumulation in double exactly as the reference does for Float = double.
 * vb200_regions_count/dim/samples/free work on both kinds of table. */
int vb200_regions_generate_single_f64(vb200_ctx* ctx, const vb200_integrand* f, const vb200_domain_f64* domain, int rule, vb200_regions** out);
int vb200_regions_generate_adaptive_f64(vb200_ctx* ctx, const vb200_integrand* f, const vb200_adaptive_params_f64* p, vb200_regions** out);
int vb200_regions_upload_f64(vb200_ctx* ctx, int dim, int rule, uint64_t count,
                             const double* rmin, const double* rmax, const double* err, const uint32_t* errdim, const double* data,
                             vb200_regions** out);
int vb200_regions_download_f64(vb200_ctx* ctx, const vb200_regions* r, double* rmin, double* rmax, double* err, uint32_t* errdim, double* data);
int vb200_regions_integrate_bins_f64(vb200_ctx* ctx, const vb200_regions* r, const vb200_domain_f64* domain, const vb200_shard* shard,
                                     double* bins, int bins_mem);
/* double-precision built-in integrands: "x2y2", "ind2", "cubic1", "poly3", "smooth_edge2", "shade4_16" */
const vb200_integrand* vb200_builtin_integrand_f64(const char* name, int exact);

/* ---- control variates + residual Monte Carlo (rows a15-a18) --------------------------------------------- */
typedef enum vb200_cv_weight {
    VB200_CV_OPTIMIZE_WEIGHT = 0,   /* cv_optimize_weight: alpha = clamp(cov,0,var)/var from the samples (weight-strategy.h:40-110) */
    VB200_CV_FIXED_WEIGHT = 1       /* cv_fixed_weight(alpha): 0 = plain importance-sampled MC ... 1 = full control variate (weight-strategy.h:7-35) */
} vb200_cv_weight;
typedef enum vb200_rr_policy {      /* Russian roulette among the regions that touch a bin (reference src/control-variates/region-russian-roulette.h) */
    VB200_RR_UNIFORM = 0,           /* rr_uniform_region  :9-28   every region equally likely */
    VB200_RR_INTEGRAL = 1,          /* rr_integral_region :30-67  probability ~ |integral of the interpolant over bin ∩ region| (floored at 1 % of the mean) */
    VB200_RR_ERROR = 2,             /* rr_error_region    :69-106 probability ~ |Region::error()| * vol(bin ∩ region)/vol(region) (same floor); nested rules only */
    VB200_RR_PDF = 3,               /* rr_pdf_region      :108-147 probability ~ integral over bin ∩ region of the shifted |interpolant| (Simpson-based rules only; factor_prob 0.01f) */
    VB200_RR_STRATIFIED = 4         /* region_stratification_uniform (region-stratification.h:9-25), the allocation of RegionsIntegratorParallelVarianceReductionOptimized
                                       (…-variance-reduction-optimized.h:120-133): spp/n samples per region of the bin, the remainder from one random start, each
                                       residual term weighted by spp / (samples of its region) */
} vb200_rr_policy;
typedef enum vb200_rs_policy {      /* where inside bin ∩ region a residual sample falls (reference src/control-variates/region-sampling.h) */
    VB200_RS_UNIFORM = 0,           /* region_sampling_uniform          :9-20   weight = volume                                              */
    VB200_RS_IMPORTANCE = 1,        /* region_sampling_importance       :22-44  position ~ |interpolant| (Simpson::sample), weight = 1/pdf   */
    VB200_RS_MIS = 2,               /* region_sampling_mis(power,cutoff):85-135 importance position or uniform position, constant MIS weight */
    VB200_RS_RUSSIAN_ROULETTE = 3   /* region_sampling_russian_roulette :46-83                                                               */
} vb200_rs_policy;
typedef struct vb200_cv_params {
    vb200_domain domain;
    vb200_shard  shard;
    uint64_t     spp;
    uint64_t     seed;
    int32_t      weight_strategy;   /* vb200_cv_weight */
    int32_t      rr_policy;         /* vb200_rr_policy (0 = the crespo2021 preset) */
    double       alpha;             /* VB200_CV_FIXED_WEIGHT only */
    int32_t      rs_policy;         /* vb200_rs_policy (0 = the crespo2021 preset); the non-uniform ones need a Simpson-based table of <= 5 dimensions
                                       and have statistical parity only (the reference inverts a cubic CDF with pow/acos/cos) */
    int32_t      reserved;
    double       rs_power, rs_cutoff; /* VB200_RS_MIS: region_sampling_mis(power = 1, cutoff = 0) */
} vb200_cv_params;

/* RegionsIntegratorParallelVarianceReduction with rr_uniform_region | rr_integral_region | rr_error_region | rr_pdf_region / cv_optimize_weight | cv_fixed_weight / region_sampling_uniform
 * (reference src/control-variates/regions-integrator-parallel-variance-reduction.h:32-109; the integrator_crespo2021
 * preset, integrator-crespo2021.h:7-22).  bins overwritten ('=').  Optional per-bin records (same memory space as
 * bins, may be NULL): nregions (uint32), approx (float, the control-variate integral). */
int vb200_cv_integrate(vb200_ctx* ctx, const vb200_integrand* f, const vb200_regions* r, const vb200_cv_params* p,
                       float* bins, int bins_mem, uint32_t* nregions, float* approx);

/* Replay: chosen [nbins_of_shard*spp] = index into the region table, samples [nbins_of_shard*spp*dim]; accumulates
 * the reference's online moments sequentially in double — bit-identical bins with an EXACT integrand. */
int vb200_cv_replay(vb200_ctx* ctx, const vb200_integrand* f, const vb200_regions* r, const vb200_cv_params* p,
                    const uint32_t* chosen, const float* samples, int mem, float* bins, int bins_mem);

/* ---- kernel launch structs (filled by the library, consumed by the thunks) ------------------------------ */
/* Chunk completion signalling of the sampling kernels (end-to-end path with VB200_HOST bins): the kernel stores its bins
 * straight into pinned, device-mapped host memory; the shard's tiles (one warp step = 32/lanes_per_bin bins) are grouped
 * into chunks of 2^chunk_shift consecutive tiles, every finished tile bumps its chunk's counter behind a system-scope fence,
 * and the warp that completes a chunk publishes `epoch` in the chunk's host-mapped flag.  The host applies the
 * reference's '+=' / '=' to a chunk as soon as its flag shows up, while the same launch is still computing the rest. */
typedef struct vb200_chunk_signal {
    uint32_t  enabled;                /* 0: no signalling (device-resident output) */
    uint32_t  chunk_shift;            /* tiles per chunk = 1 << chunk_shift */
    uint32_t* done;                   /* device [chunks], zero at launch: finished tiles per chunk (the warp that completes a chunk zeroes its counter again) */
    uint32_t* flag;                   /* host-mapped [chunks]: set to epoch when every tile of the chunk has been stored */
    uint32_t  epoch;                  /* changes with every call, so the flags never need clearing */
    uint32_t  reserved;
} vb200_chunk_signal;

typedef struct vb200_mc_launch {
    vb200_domain domain;
    uint64_t bin_begin, bin_end;      /* shard */
    uint64_t nbins_total;
    uint32_t spp;
    uint32_t lanes_per_bin;           /* power of two, 1..32 */
    uint32_t key0, key1;              /* Philox key */
    int32_t  flavor;
    int32_t  accumulate;              /* 1: out[b] = float(double(out[b]) + v) ; 0: out[b] = v */
    double   factor;                  /* flavor 0: vol(range)/spp */
    float*   out;                     /* device, base of full grid */
    float*   sum_f;                   /* device or NULL, base of shard */
    float*   sum_f2;
    int32_t  grid_hint;               /* CTAs to launch (0 = let the thunk size it from occupancy) */
    int32_t  narrow_binned;           /* 1: coordinates of the binned dimensions carry 16 random bits (every binned dimension of the
                                       * WHOLE grid has >= 256 bins, so the lattice along it still has >= 2^24 points); 0: 24 bits */
    unsigned long long* tile_counter; /* device, two words, zero at launch: [0] tile tickets, [1] warps that are through — the kernel zeroes both again when
                                       * its last warp leaves, so the driver clears them once per context, not once per call */
    vb200_chunk_signal signal;
    int32_t  rng;                     /* VB200_RNG_XOSHIRO: xoshiro128++ stream per (bin, lane sub-stream) seeded by Philox; VB200_RNG_PHILOX: Philox4x32-10 per draw group */
    int32_t  reserved;
} vb200_mc_launch;

typedef struct vb200_replay_launch {
    vb200_domain domain;
    uint64_t bin_begin, bin_end, nbins_total;
    uint32_t spp; int32_t flavor;
    double   factor;
    const float* samples;             /* device, [nbins_of_shard][spp][dim] */
    float*   out;                     /* device, base of full grid; read-modify-write for flavor 0 */
} vb200_replay_launch;

typedef struct vb200_walk_launch {
    vb200_domain domain;              /* dim = explicit range entries */
    uint64_t bin_begin, bin_end, nbins_total;
    uint32_t spp; uint32_t lanes_per_bin;
    uint32_t key0, key1;
    int32_t  accumulate; int32_t grid_hint;
    double   factor;
    float*   out; float* sum_f; float* sum_f2;
    unsigned long long* tile_counter; /* device, zeroed by the driver before the launch */
    int32_t  flavor; int32_t reserved;
    vb200_chunk_signal signal;
} vb200_walk_launch;

typedef struct vb200_walk_replay_launch {
    vb200_domain domain;
    uint64_t bin_begin, bin_end, nbins_total;
    uint32_t spp; int32_t flavor;
    double   factor;
    const uint64_t* offsets; const float* elems;
    float*   out;
    int32_t* error_flag;              /* device int, set to 1 when a path reads past its recorded length */
} vb200_walk_replay_launch;

typedef struct vb200_eval_launch {
    uint64_t n;
    int32_t  dim; int32_t f64;        /* f64 != 0: points/values are doubles (VB200_INTEGRAND_F64 integrands) */
    const void* points;               /* device, SoA: points[d*n + i] */
    void*    values;                  /* device, [n] */
} vb200_eval_launch;

typedef struct vb200_greedy_launch {
    int32_t  dim, rule, heuristic, metric;
    double   size_weight;
    uint64_t iterations;
    uint64_t capacity;                /* region slots: 2*iterations+1 (a split retires the parent slot and opens two) */
    /* working set of the persistent kernel, region-major so that one region is a few contiguous lines; float or double per f64: */
    void*    range;                   /* device [capacity][2*dim]: min[0..dim), max[0..dim) */
    void*    data;                    /* device [capacity][S^dim], the reference's multiarray order (dim 0 fastest) */
    void*    err;                     /* device [capacity] heuristic value in the table's scalar type (float-key runs) */
    void*    heap;                    /* device [iterations+2]: binary heap entries — 8 bytes (id | dim<<28) << 32 | float bits of the key, or, when the
                                         keys are doubles (f64 tables, VB200_HEURISTIC_MIXED), 16 bytes {double bits, id | dim<<28}; array order = the
                                         reference's output order (regions-generator-adaptive-heap.h:44) */
    uint64_t* heap_size;              /* device scalar: entries in the heap when the kernel ends (iterations+1) */
    float    range_min[VB200_MAX_DIM], range_max[VB200_MAX_DIM];
    /* ABI 3: */
    int32_t  f64;                     /* 1: double table (range_min64/range_max64, functor over std::array<double,dim>) */
    int32_t  metric_rest, mixed_dimension, reserved;
    double   mixed_bins_weight, mixed_threshold_bins, mixed_threshold_rest, mixed_error_increase;
    double   range_min64[VB200_MAX_DIM], range_max64[VB200_MAX_DIM];
    void*    key64;                   /* device [capacity] double heuristic values (double-key runs), else NULL */
} vb200_greedy_launch;

typedef struct vb200_scatter_launch {
    vb200_domain domain;
    uint64_t sample_begin, sample_end, nbins_total;
    uint32_t key0, key1;
    double   factor;
    float*   out;
    int32_t  grid_hint; int32_t reserved;
} vb200_scatter_launch;

#ifdef __cplusplus
}
#endif
#endif /* VILTRUM_B200_H */
