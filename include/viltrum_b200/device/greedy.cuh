// K5-K7, exact mode — greedy max-error refinement with batch size 1, replacing the reference's
//   RegionsGeneratorAdaptiveHeap::generate    src/nested/regions-generator-adaptive-heap.h:18-45
//   Region ctor / multiarray::fill            src/newton-cotes/region.h:62-68, src/multiarray/fill.h:45-72
//   Region::split / detail::split             src/newton-cotes/region.h:345-359, src/multiarray/split.h:13-49
//   Region::error, error_heuristic_*          src/newton-cotes/region.h:387-411, src/nested/error-heuristic.h:10-46
//   std::push_heap / std::pop_heap            libstdc++ bits/stl_heap.h:135-267 (tie order is decided by these mechanics)
// The loop is inherently serial (every iteration depends on which region the previous one made the maximum), so a launch
// per iteration cannot work: ONE persistent CTA runs all iterations.  Inside an iteration the work is spread over the CTA:
//   warp 0 / lane 0  pops the heap (its top levels live in shared memory, the rest in L2) while
//   the other warps  fetch the parent region, evaluate the integrand at the (S-1)*S^(D-1) new points of the split and
//                    build both children in shared memory;
//   one warp per (child, dimension) evaluates the nested-rule error along that dimension and folds it over the others;
//   then lane 0 pushes the two children.
// With an EXACT integrand (--fmad=false) the region list — ranges, samples, errors, split dimensions and ORDER — is
// bit-identical to the reference's (tests/test_gpu_regions.py).
#pragma once
#include <array>
#include <type_traits>
#include <cuda_runtime.h>
#include "../../viltrum_b200.h"
#include "rules.cuh"

namespace viltrum { namespace b200 { namespace device {

constexpr int GREEDY_THREADS = 256;
constexpr unsigned GREEDY_ID_MASK = 0x0fffffffu;

// Heap entries.  Float keys (error_heuristic_default / _size over Range<float>): one 64-bit word, (id | dim << 28) << 32 | key bits.
// Double keys (Range<double>, and error_heuristic_mixed, whose key is a double upstream, error-heuristic.h:73-96): 16 bytes.
struct GreedyEntry32 {
    typedef unsigned long long type; typedef float key_type;
    __device__ __forceinline__ static type make(unsigned id_dim, float k) { return (static_cast<unsigned long long>(id_dim) << 32) | __float_as_uint(k); }
    __device__ __forceinline__ static float key(type e) { return __uint_as_float(unsigned(e)); }
    __device__ __forceinline__ static unsigned id_dim(type e) { return unsigned(e >> 32); }
};
struct GreedyEntry64 {
    typedef ulonglong2 type; typedef double key_type;
    __device__ __forceinline__ static type make(unsigned id_dim, double k) { return make_ulonglong2(static_cast<unsigned long long>(__double_as_longlong(k)), id_dim); }
    __device__ __forceinline__ static double key(type e) { return __longlong_as_double(static_cast<long long>(e.x)); }
    __device__ __forceinline__ static unsigned id_dim(type e) { return unsigned(e.y); }
};

// The heap is operated by ONE thread and lives in that thread's registers (pointers, cache size, count): kept in shared memory, every
// get/set re-read those fields behind each store — ~290 cycles per heap level, 5200 of the 10200 cycles of a BASELINE-config-3 iteration
// (profiles/greedy_phases_r2.txt).  Indices are 32-bit (capacity <= 2^28).
template<class E = GreedyEntry32>
struct GreedyHeapT {
    typedef typename E::type entry; typedef typename E::key_type key_type;
    entry* g;      // global entries
    entry* s;      // shared-memory cache of entries [0, cached)
    unsigned cached;
    unsigned n;
    __device__ __forceinline__ entry get(unsigned i) const { return i < cached ? s[i] : g[i]; }
    __device__ __forceinline__ void set(unsigned i, entry v) { if (i < cached) s[i] = v; else g[i] = v; }
    __device__ __forceinline__ static key_type key(entry e) { return E::key(e); }

    // libstdc++ __push_heap (stl_heap.h:135-148), comparator a.err < b.err
    __device__ __forceinline__ void sift_up(unsigned hole, entry value) {
        const key_type vk = key(value);
        while (hole > 0) {
            const unsigned parent = (hole - 1u) >> 1;
            const entry pe = get(parent);
            if (!(key(pe) < vk)) break;
            set(hole, pe); hole = parent;
        }
        set(hole, value);
    }
    __device__ __forceinline__ void push(entry value) { ++n; sift_up(n - 1u, value); }
    // libstdc++ pop_heap -> __pop_heap -> __adjust_heap (stl_heap.h:224-267) followed by the caller's pop_back
    __device__ __forceinline__ void pop() {
        if (n > 1u) {
            const entry value = get(n - 1u);
            const unsigned len = n - 1u;
            unsigned hole = 0, child = 0;
            while (child < (len - 1u) / 2u) {
                child = 2u * (child + 1u);
                entry ce = get(child); const entry le = get(child - 1u);
                if (key(ce) < key(le)) { --child; ce = le; }
                set(hole, ce); hole = child;
            }
            if ((len & 1u) == 0u && child == (len - 2u) / 2u) { child = 2u * (child + 1u); set(hole, get(child - 1u)); hole = child - 1u; }
            sift_up(hole, value);
        }
        --n;
    }
};
typedef GreedyHeapT<GreedyEntry32> GreedyHeap;

template<int SH, int SL, int DIM>
struct GreedyShape {
    static constexpr int pow_(int b, int e) { return e == 0 ? 1 : b * pow_(b, e - 1); }
    static constexpr int SD = pow_(SH, DIM);          // samples per region
    static constexpr int L = pow_(SH, DIM - 1);       // lines along one dimension
    static constexpr int WIDE = (2 * SH - 1) * L;     // samples of the two children side by side
};

// normalised grid coordinate -> point, with the PARENT's range (region.h:40-46): x = Float(p*(max-min) + min)
template<class T>
__device__ __forceinline__ T grid_coord(double p, T lo, T hi) {
    return rules::from_double<T>(rules::da(rules::dm(p, double(rules::sub(hi, lo))), double(lo)));
}

// error of one region along `dim` (region.h:387-393): per line metric(high,low), folded over the other dims with the high
// rule, times the volume.  One warp; `work` holds L values.
template<int SH, int SL, int DIM, class T = float>
__device__ T region_error_warp(const T* data, T volume, int dim, bool relative, T* work, unsigned lane) {
    using Sh = GreedyShape<SH, SL, DIM>;
    int inner = 1; for (int i = 0; i < dim; ++i) inner *= SH;
    for (int o = lane; o < Sh::L; o += 32) {
        const int lo = o % inner, hi = o / inner;
        T line[SH];
#pragma unroll
        for (int e = 0; e < SH; ++e) line[e] = data[lo + e * inner + hi * inner * SH];
        work[o] = rules::line_error<SH, SL, T>(relative, line);
    }
    __syncwarp();
    // fold_all(high rule): fold dimension 0 of the remaining array until one value is left (fold.h:87-108)
    for (int n = Sh::L / SH; n >= 1; n /= SH) {
        T v[(Sh::L / SH + 31) / 32 > 0 ? (Sh::L / SH + 31) / 32 : 1];
        int c = 0;
        for (int o = lane; o < n; o += 32, ++c) v[c] = rules::apply<SH, T>(work + o * SH);
        __syncwarp();
        c = 0;
        for (int o = lane; o < n; o += 32, ++c) work[o] = v[c];
        __syncwarp();
        if (n == 1) break;
    }
    return rules::mul(volume, work[0]);
}

// error_heuristic_size (error-heuristic.h:29-46) / error_heuristic_default -> max_error_dimension (region.h:401-411); key type = Float
template<int DIM, class T = float>
__device__ void heuristic_pick(const T* E, const T* rng /* min[DIM], max[DIM] */, int heuristic, double size_weight, T* out_err, unsigned* out_dim) {
    const double min_size = 1.e-37;
    if (heuristic == VB200_HEURISTIC_SIZE) {
        T max_err = E[0];
        const T w0 = rules::sub(rng[DIM], rng[0]);
        if (double(w0) < min_size || isnan(w0)) max_err = T(0);
        else max_err = rules::from_double<T>(rules::da(double(max_err), rules::dm(size_weight, double(rules::absv(w0)))));
        unsigned max_dim = 0;
        for (int d = 1; d < DIM; ++d) {
            const T w = rules::sub(rng[DIM + d], rng[d]);
            T err = rules::from_double<T>(rules::da(double(E[d]), rules::dm(size_weight, double(rules::absv(w)))));
            if (double(w) < min_size) err = T(0);
            if (err >= max_err) { max_err = err; max_dim = unsigned(d); }
        }
        *out_err = max_err; *out_dim = max_dim;
    } else {
        T max_err = T(0); unsigned max_dim = 0;
        for (int d = 0; d < DIM; ++d) if (E[d] > max_err) { max_err = E[d]; max_dim = unsigned(d); }
        *out_err = max_err; *out_dim = max_dim;
    }
}
// error_heuristic_mixed (error-heuristic.h:49-98): the key is a DOUBLE upstream (Float * double + double + double * Float).  E[d] holds the error
// along d under the metric that dimension takes (bins metric for d < dimension and for d = 0, rest metric beyond).
struct MixedParams { int dimension; double bins_weight, size_weight, size_threshold_bins, size_threshold_rest, error_increase_factor; };
template<int DIM, class T>
__device__ void heuristic_pick_mixed(const T* E, const T* rng, const MixedParams& m, double* out_err, unsigned* out_dim) {
    double size_bins = 1.0, size_rest = 1.0;
    const int nb = m.dimension < DIM ? m.dimension : DIM;
    for (int d = 0; d < nb; ++d) size_bins = rules::dm(size_bins, double(rules::absv(rules::sub(rng[DIM + d], rng[d]))));
    for (int d = m.dimension; d < DIM; ++d) size_rest = rules::dm(size_rest, double(rules::absv(rules::sub(rng[DIM + d], rng[d]))));
    double add_bins = m.error_increase_factor, add_rest = m.error_increase_factor;
    if (size_bins < m.size_threshold_bins) add_bins = 0.0;
    if (size_rest < m.size_threshold_rest) add_rest = 0.0;
    if (isnan(size_bins)) add_bins = 0.0;
    if (isnan(size_rest)) add_rest = 0.0;
    double max_err = rules::da(rules::da(rules::dm(double(E[0]), m.bins_weight), add_bins), rules::dm(m.size_weight, double(rules::sub(rng[DIM], rng[0]))));
    unsigned max_dim = 0;
    for (int d = 1; d < DIM; ++d) {
        const double sz = rules::dm(m.size_weight, double(rules::sub(rng[DIM + d], rng[d])));
        const double err = d < m.dimension ? rules::da(rules::da(rules::dm(double(E[d]), m.bins_weight), add_bins), sz)
                                           : rules::da(rules::da(double(E[d]), add_rest), sz);
        if (err >= max_err) { max_err = err; max_dim = unsigned(d); }
    }
    *out_err = max_err; *out_dim = max_dim;
}

#ifdef VB200_GREEDY_TIMING      // experiment builds (profiles/exp/greedy_phases.cu): cycles thread 0 spends between the barriers of an iteration
__device__ unsigned long long vb200_greedy_clock[8];
#define VB200_GT(k) do { if (tid == 0) { const long long now_ = clock64(); vb200_greedy_clock[k] += (unsigned long long)(now_ - t_prev_); t_prev_ = now_; } } while (0)
#else
#define VB200_GT(k) do {} while (0)
#endif
// T = float or double (the Float of the range); MIXED selects error_heuristic_mixed (double keys).  Keys are doubles whenever T is double or MIXED.
template<class F, int DIM, int SH, int SL, bool EXACT, class T = float, bool MIXED = false>
__global__ void __launch_bounds__(GREEDY_THREADS, 1)
greedy_kernel(const F f, const vb200_greedy_launch a, const int heap_cached) {
    using Sh = GreedyShape<SH, SL, DIM>;
    constexpr bool KEY64 = MIXED || sizeof(T) == 8;
    using E = typename std::conditional<KEY64, GreedyEntry64, GreedyEntry32>::type;
    using Heap = GreedyHeapT<E>;
    using Key = typename E::key_type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typename E::type* s_heap = reinterpret_cast<typename E::type*>(smem_raw);
    T* s_parent = reinterpret_cast<T*>(s_heap + heap_cached);           // [SD]
    T* s_child = s_parent + Sh::SD;                                     // [2][SD]
    T* s_work = s_child + 2 * Sh::SD;                                   // [2*DIM][L]
    T* s_prange = s_work + 2 * DIM * Sh::L;                             // [2*DIM]
    T* s_crange = s_prange + 2 * DIM;                                   // [2][2*DIM]
    T* s_E = s_crange + 4 * DIM;                                        // [2][DIM]
    T* s_vol = s_E + 2 * DIM;                                           // [2]
    __shared__ unsigned s_top_id, s_top_dim;
    __shared__ unsigned s_pick_dim[2];
    __shared__ double s_pick_key[2];          // heap keys of the two children (exactly representable: Key is float or double)
    Heap heap;                                // thread 0's registers

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, nwarps = GREEDY_THREADS / 32;
    const bool relative = a.metric == VB200_METRIC_RELATIVE, relative_rest = a.metric_rest == VB200_METRIC_RELATIVE;
    const MixedParams mixed{a.mixed_dimension, a.mixed_bins_weight, a.size_weight, a.mixed_threshold_bins, a.mixed_threshold_rest, a.mixed_error_increase};
    // metric of dimension d: error_heuristic_mixed uses the bins metric for d = 0 and d < dimension, the rest metric beyond (error-heuristic.h:81-88)
    auto rel_of = [&] (int d) -> bool { return MIXED ? ((d == 0 || d < mixed.dimension) ? relative : relative_rest) : relative; };
    auto pick = [&] (const T* Ev, const T* rng, Key* key, unsigned* dim) {
        if constexpr (MIXED) heuristic_pick_mixed<DIM, T>(Ev, rng, mixed, key, dim);
        else { T e; heuristic_pick<DIM, T>(Ev, rng, a.heuristic, a.size_weight, &e, dim); *key = Key(e); }
    };
    // i / m for the grid positions: m = S-1 or 2(S-1) is a power of two for S = 3, 5, so the division is an exact multiplication
    auto frac = [] (int i, int m) -> double { return ((m & (m - 1)) == 0) ? rules::dm(double(i), 1.0 / double(m)) : rules::dd(double(i), double(m)); };
    T* g_range = static_cast<T*>(a.range); T* g_data = static_cast<T*>(a.data); T* g_err = static_cast<T*>(a.err);
    const T* rmin_ = reinterpret_cast<const T*>(sizeof(T) == 8 ? static_cast<const void*>(a.range_min64) : static_cast<const void*>(a.range_min));
    const T* rmax_ = reinterpret_cast<const T*>(sizeof(T) == 8 ? static_cast<const void*>(a.range_max64) : static_cast<const void*>(a.range_max));

    // ---- initial region over the whole range (regions-generator-adaptive-heap.h:27-31) ----
    heap.g = static_cast<typename E::type*>(a.heap); heap.s = s_heap; heap.cached = unsigned(heap_cached); heap.n = 0;
    if (tid < 2 * DIM) s_crange[tid] = tid < DIM ? rmin_[tid] : rmax_[tid - DIM];
    __syncthreads();
    for (int k = tid; k < Sh::SD; k += GREEDY_THREADS) {
        std::array<T, DIM> x; int t = k;
#pragma unroll
        for (int d = 0; d < DIM; ++d) { x[d] = grid_coord<T>(frac(t % SH, SH - 1), s_crange[d], s_crange[DIM + d]); t /= SH; }
        s_child[k] = f(x);
    }
    if (tid == 0) { T v = T(1); for (int d = 0; d < DIM; ++d) v = rules::mul(v, rules::sub(s_crange[DIM + d], s_crange[d])); s_vol[0] = v; }
    __syncthreads();
    for (int d = warp; d < DIM; d += nwarps) {
        const T e = region_error_warp<SH, SL, DIM, T>(s_child, s_vol[0], d, rel_of(d), s_work + d * Sh::L, lane);
        if (lane == 0) s_E[d] = e;
    }
    __syncthreads();
    if (tid == 0) {
        Key err; unsigned dim; pick(s_E, s_crange, &err, &dim);
        if constexpr (KEY64) static_cast<double*>(a.key64)[0] = double(err); else g_err[0] = T(err);
        heap.push(E::make(0u | (dim << 28), err));
        s_top_id = 0u; s_top_dim = dim;
    }
    for (int k = tid; k < Sh::SD; k += GREEDY_THREADS) g_data[k] = s_child[k];
    if (tid < 2 * DIM) g_range[tid] = s_crange[tid];

    // ---- iterations ----
    // Two CTA barriers per iteration.  Between B1 and B2 warp 0 pops the heap (thread 0; the pop does not depend on the split: heap.front()
    // was copied first, :33-35) WHILE warps 1..7 fetch the parent, evaluate the split, fold both children's errors and pick their
    // heuristics (barrier 1 of 224 threads between those steps); between B2 and B1 thread 0 pushes the two children and publishes the
    // next top while the other threads store the children to the region arrays.
    constexpr int WORKERS = GREEDY_THREADS - 32;
    auto worker_sync = [] () { asm volatile("bar.sync 1, %0;" :: "n"(WORKERS) : "memory"); };
    unsigned long long next_slot = 1;
#ifdef VB200_GREEDY_TIMING
    long long t_prev_ = clock64();
#endif
    for (unsigned long long it = 0; it < a.iterations; ++it) {
        __syncthreads();                                                   // B1: top published, previous children stored
        VB200_GT(0);
        const unsigned top = s_top_id; const int dim = int(s_top_dim);
        if (warp == 0) {
            if (lane == 0) heap.pop();
            VB200_GT(1);
        } else {
            const int wt = int(tid) - 32;
            for (int k = wt; k < Sh::SD; k += WORKERS) s_parent[k] = g_data[static_cast<unsigned long long>(top) * Sh::SD + k];
            if (wt < 2 * DIM) s_prange[wt] = g_range[static_cast<unsigned long long>(top) * (2 * DIM) + wt];
            worker_sync();
            // split along `dim` (split.h:13-49): the (2S-1)-wide array; even positions are the parent's samples, odd ones new evaluations
            int inner = 1; for (int i = 0; i < dim; ++i) inner *= SH;
            for (int item = wt; item < Sh::WIDE; item += WORKERS) {
                const int i = item / Sh::L, o = item % Sh::L;            // position along `dim` (0..2S-2), index over the other dims
                const int lo = o % inner, hi = o / inner;
                T v;
                if ((i & 1) == 0) v = s_parent[lo + (i / 2) * inner + hi * inner * SH];
                else {
                    std::array<T, DIM> x; int t = o;
#pragma unroll
                    for (int d = 0; d < DIM; ++d) {
                        double p;
                        if (d == dim) p = frac(i, 2 * (SH - 1));
                        else { p = frac(t % SH, SH - 1); t /= SH; }
                        x[d] = grid_coord<T>(p, s_prange[d], s_prange[DIM + d]);
                    }
                    v = f(x);
                }
                if (i <= SH - 1) s_child[lo + i * inner + hi * inner * SH] = v;
                if (i >= SH - 1) s_child[Sh::SD + lo + (i - (SH - 1)) * inner + hi * inner * SH] = v;
            }
            // child ranges (region.h:349-357): d = (max-min)/Float(2); child 0 = [min, min+d*1], child 1 = [min+d*1, max]; the last worker warp
            if (wt >= WORKERS - 2) {
                const int c = wt - (WORKERS - 2);
                const T pmin = s_prange[dim], pmax = s_prange[DIM + dim];
                const T mid = rules::add(pmin, rules::mul(rules::quo(rules::sub(pmax, pmin), T(2)), T(1)));
                T* cr = s_crange + c * 2 * DIM;
                for (int d = 0; d < 2 * DIM; ++d) cr[d] = s_prange[d];
                if (c == 0) cr[DIM + dim] = mid; else cr[dim] = mid;
                T v = T(1); for (int d = 0; d < DIM; ++d) v = rules::mul(v, rules::sub(cr[DIM + d], cr[d]));
                s_vol[c] = v;
            }
            worker_sync();
            // nested-rule error of both children along every dimension: one worker warp per (child, dimension)
            for (int job = int(warp) - 1; job < 2 * DIM; job += int(nwarps) - 1) {
                const int c = job / DIM, d = job % DIM;
                const T e = region_error_warp<SH, SL, DIM, T>(s_child + c * Sh::SD, s_vol[c], d, rel_of(d), s_work + job * Sh::L, lane);
                if (lane == 0) s_E[c * DIM + d] = e;
            }
            worker_sync();
            if (wt < 2) {                                                   // the two children's heuristics side by side (:36-40)
                Key err; unsigned d; pick(s_E + wt * DIM, s_crange + wt * 2 * DIM, &err, &d);
                s_pick_key[wt] = double(err); s_pick_dim[wt] = d;
            }
        }
        __syncthreads();                                                   // B2: popped; children, errors and picks ready
        VB200_GT(2);
        if (tid == 0) {
            for (int c = 0; c < 2; ++c) {
                const Key err = Key(s_pick_key[c]); const unsigned d = s_pick_dim[c];
                heap.push(E::make(unsigned(next_slot + c) | (d << 28), err));
            }
            const unsigned idd = E::id_dim(heap.get(0)); s_top_id = idd & GREEDY_ID_MASK; s_top_dim = idd >> 28;
            VB200_GT(3);
        } else {
            // store the children (slots next_slot, next_slot+1)
            for (int k = int(tid) - 1; k < 2 * Sh::SD; k += GREEDY_THREADS - 1) g_data[next_slot * Sh::SD + k] = s_child[k];
            if (tid - 1 < 4 * DIM) g_range[next_slot * (2 * DIM) + (tid - 1)] = s_crange[tid - 1];
            if (tid - 1 < 2) { if constexpr (KEY64) static_cast<double*>(a.key64)[next_slot + (tid - 1)] = s_pick_key[tid - 1]; else g_err[next_slot + (tid - 1)] = T(s_pick_key[tid - 1]); }
        }
        next_slot += 2;
    }
    __syncthreads();
    // flush the cached top of the heap
    if (tid == 0) { s_top_id = heap.n; *a.heap_size = static_cast<uint64_t>(heap.n); }
    __syncthreads();
    const unsigned n = s_top_id;
    for (unsigned i = tid; i < n && i < unsigned(heap_cached); i += GREEDY_THREADS) static_cast<typename E::type*>(a.heap)[i] = s_heap[i];
}

template<class F, int DIM, int SH, int SL, bool EXACT, class T, bool MIXED>
inline int launch_greedy_rule(const F& f, const vb200_greedy_launch& a, cudaStream_t st) {
    using Sh = GreedyShape<SH, SL, DIM>;
    constexpr bool KEY64 = MIXED || sizeof(T) == 8;
    constexpr size_t ENTRY = KEY64 ? 16 : 8;
    if (a.capacity > GREEDY_ID_MASK) return int(cudaErrorInvalidValue);
    auto k = greedy_kernel<F, DIM, SH, SL, EXACT, T, MIXED>;
    const size_t fixed = sizeof(T) * size_t(3 * Sh::SD + 2 * DIM * Sh::L + 2 * DIM + 4 * DIM + 2 * DIM + 2) + 64;
    int dev = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (size_t(max_smem) < fixed + 1024) return int(cudaErrorInvalidConfiguration);     // region too large for the one-CTA working set
    // cache as many complete top levels of the heap as fit: 2^k - 1 entries
    size_t room = size_t(max_smem) - fixed - 1024;
    int cached = 1; while (size_t(2 * cached + 1) * ENTRY <= room && cached < (1 << 15)) cached = 2 * cached + 1;
    const size_t smem = size_t(cached) * ENTRY + fixed;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return int(e);
    k<<<1, GREEDY_THREADS, smem, st>>>(f, a, cached);
    return int(cudaGetLastError());
}

template<class F, int DIM, bool EXACT, class T = float>
inline int launch_greedy(const F& f, const vb200_greedy_launch& a, cudaStream_t st) {
    const bool mixed = a.heuristic == VB200_HEURISTIC_MIXED;
    if constexpr (DIM <= 6) {
        if (a.rule == VB200_RULE_SIMPSON_TRAPEZOIDAL) return mixed ? launch_greedy_rule<F, DIM, 3, 2, EXACT, T, true>(f, a, st) : launch_greedy_rule<F, DIM, 3, 2, EXACT, T, false>(f, a, st);
    }
    if constexpr (DIM <= 5 && sizeof(T) == 4) {
        if (a.rule == VB200_RULE_BOOLE_SIMPSON) return mixed ? launch_greedy_rule<F, DIM, 5, 3, EXACT, T, true>(f, a, st) : launch_greedy_rule<F, DIM, 5, 3, EXACT, T, false>(f, a, st);
    }
    if constexpr (DIM <= 4 && sizeof(T) == 8) {      // doubles: twice the shared memory per sample
        if (a.rule == VB200_RULE_BOOLE_SIMPSON) return mixed ? launch_greedy_rule<F, DIM, 5, 3, EXACT, T, true>(f, a, st) : launch_greedy_rule<F, DIM, 5, 3, EXACT, T, false>(f, a, st);
    }
    return int(cudaErrorNotSupported);
}

}}} // namespace viltrum::b200::device
