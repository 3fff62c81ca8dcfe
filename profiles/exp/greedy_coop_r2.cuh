// EXPERIMENT, NOT SHIPPED (round 2): a cooperative variant of include/viltrum_b200/device/greedy.cuh, kept for the record.
//   heap warp: subtree prefetch in pop (126 entries per L2 round trip), one level of look-ahead in the walk, path-parallel pushes over
//   prefetched ancestors, prefetch of the likeliest next parent; one worker warp (64-thread CTA) for regions of <= 2 dimensions.
// Result (profiles/greedy_coop_r2.txt): region lists stay bit-identical (181 exact-mode tests passed with it), but 2.70-2.85 us per iteration
// against 2.56 us of the shipped kernel.  The phase counters show why: the serial side runs at ~5 cycles per dependent instruction on a single
// warp, a heap level is ~25 such instructions (120-127 cycles) whatever the memory latency behind it, and the cooperative bookkeeping adds
// instructions to the same serial stream.  Getting under the CPU's 0.83 us would take <= ~450 serial instructions per iteration.
// K5-K7, exact mode — greedy max-error refinement with batch size 1, replacing the reference's
//   RegionsGeneratorAdaptiveHeap::generate    src/nested/regions-generator-adaptive-heap.h:18-45
//   Region ctor / multiarray::fill            src/newton-cotes/region.h:62-68, src/multiarray/fill.h:45-72
//   Region::split / detail::split             src/newton-cotes/region.h:345-359, src/multiarray/split.h:13-49
//   Region::error, error_heuristic_*          src/newton-cotes/region.h:387-411, src/nested/error-heuristic.h:10-46
//   std::push_heap / std::pop_heap            libstdc++ bits/stl_heap.h:135-267 (tie order is decided by these mechanics)
// The loop is inherently serial (every iteration depends on which region the previous one made the maximum), so a launch
// per iteration cannot work: ONE persistent CTA runs all iterations.  Inside an iteration the work is spread over the CTA:
//   warp 0           operates the heap (its top levels live in shared memory, the rest in L2): pop, then the two pushes, while
//   the other warps  take the parent region (from shared memory: it is either the region warp 0 prefetched or one of the last two
//                    children), evaluate the integrand at the (S-1)*S^(D-1) new points of the split, build both children and their
//                    nested-rule errors along every dimension.  Regions of up to two dimensions need ONE worker warp (64-thread CTA).
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

#ifdef VB200_GREEDY_TIMING      // experiment builds (profiles/exp/greedy_phases.cu): cycles thread 0 spends in each phase of an iteration
__device__ unsigned long long vb200_greedy_clock[8];
__shared__ unsigned long long vb200_gt_acc[8];      // accumulated in shared memory: a global read-modify-write per phase costs more than most phases
__shared__ long long vb200_gt_prev;
#define VB200_GT_(who, k) do { if ((who) == 0) { const long long now_ = clock64(); vb200_gt_acc[k] += (unsigned long long)(now_ - vb200_gt_prev); vb200_gt_prev = now_; } } while (0)
#define VB200_GT(k) VB200_GT_(tid, k)
#define VB200_GTL(k) VB200_GT_(lane, k)      /* inside the heap warp's methods */
#else
#define VB200_GT(k) do {} while (0)
#define VB200_GTL(k) do {} while (0)
#endif
// The heap is operated by ONE WARP whose 32 lanes hold identical copies of the bookkeeping (pointers, cache size, count) in registers and
// execute the serial parts redundantly (shared-memory reads broadcast, identical stores merge): no lane ever waits for another one to
// tell it where the hole went.  Indices are 32-bit (capacity <= 2^28).  The top `cached` = 2^k - 1 entries (complete levels) live in shared
// memory, the rest in global memory (L2).  What the warp adds over one thread (profiles/greedy_phases_r2.txt: a one-thread pop is a chain of
// ~20 dependent levels, the lower six at L2 latency = 2950-3500 cycles at 10^6 entries, and two one-thread pushes cost 1000-1100):
//   pop    below the cached levels the lanes fetch the WHOLE subtree under the hole, six levels = 126 entries, in one round trip into a
//          shared scratch area, and the walk continues at shared-memory latency;
//   push   the ancestors of a new leaf are known before its key is (positions (p+1 >> k) - 1): the lanes fetch both leaves' paths right
//          after the pop — while the other warps are still computing the children — and once the keys arrive a ballot finds where
//          each sift-up stops; every ancestor that moves is stored by its own lane.
// Mechanics and tie order are libstdc++'s (bits/stl_heap.h:135-267): __adjust_heap walks to the bottom, then __push_heap sifts the former
// last element up; push_heap sifts up while parent < value (strict).
constexpr int GREEDY_PF_LEVELS = 6, GREEDY_PF_ENTRIES = 128;
template<class E = GreedyEntry32>
struct GreedyHeapWarp {
    typedef typename E::type entry; typedef typename E::key_type key_type;
    entry* g;        // global entries
    entry* s;        // shared-memory cache of entries [0, cached)
    entry* scratch;  // shared, GREEDY_PF_ENTRIES entries: the subtree fetched by pop
    unsigned cached, n;
    struct Paths { entry a, b; unsigned ia, ib; bool va, vb; };      // this lane's ancestor on the path of either new leaf
    __device__ __forceinline__ entry get(unsigned i) const { return i < cached ? s[i] : g[i]; }
    __device__ __forceinline__ void set(unsigned i, entry v) { if (i < cached) s[i] = v; else g[i] = v; }
    __device__ __forceinline__ static key_type key(entry e) { return E::key(e); }
    __device__ __forceinline__ static unsigned long long shfl1(unsigned long long v, unsigned src) { return __shfl_sync(0xffffffffu, v, int(src & 31u)); }
    __device__ __forceinline__ static ulonglong2 shfl1(ulonglong2 v, unsigned src) { return make_ulonglong2(shfl1(v.x, src), shfl1(v.y, src)); }

    // libstdc++ __push_heap (stl_heap.h:135-148), comparator a.err < b.err; every lane runs it (rare path of pop)
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
    // __adjust_heap's descent (stl_heap.h:230-240: the hole moves to its larger child, the right one on ties) over a heap-layout array in shared
    // memory — the cached top of the heap (DEEP = false: local index = heap index, moves are stored into the array) or the fetched subtree
    // under the hole (DEEP = true: the hole is the absent root, moves go to the heap itself).  One level of look-ahead: the hole's children are in
    // registers and the four grandchildren were requested a level earlier, so a level costs a compare and a select instead of a dependent
    // address-load-compare-branch round (~120 cycles measured).  Slots past the end of the heap may be loaded (they are inside the array) but
    // never selected: a pair is only used once its parent is a hole below `limit`.  Walks until the hole reaches `limit` or the array's last level.
    template<bool DEEP>
    __device__ __forceinline__ void walk(entry* arr, unsigned size, unsigned& hole, unsigned limit, entry& moved, bool& any) {
        if (!(hole < limit) || size < 7u) return;
        unsigned q = 0;
        entry eL = arr[1], eR = arr[2];
        entry gLL = arr[3], gLR = arr[4], gRL = arr[5], gRR = arr[6];
        while (true) {
            const bool left = key(eR) < key(eL);
            const unsigned cq = 2u * q + (left ? 1u : 2u), c = 2u * hole + (left ? 1u : 2u);
            const bool more = 4u * cq + 6u < size;
            entry nLL = gLL, nLR = gLR, nRL = gRL, nRR = gRR;
            if (more) { nLL = arr[4u * cq + 3u]; nLR = arr[4u * cq + 4u]; nRL = arr[4u * cq + 5u]; nRR = arr[4u * cq + 6u]; }      // requested before the old ones are consumed
            const entry w = left ? eL : eR;
            if (DEEP) set(hole, w); else arr[q] = w;
            hole = c; q = cq; moved = w; any = true;
            eL = left ? gLL : gRL; eR = left ? gLR : gRR;
            if (!(hole < limit)) return;
            if (!more) break;
            gLL = nLL; gLR = nLR; gRL = nRL; gRR = nRR;
        }
        if (2u * q + 2u < size) {                           // the array's last level: children in registers, no grandchildren
            const bool left = key(eR) < key(eL);
            const unsigned c = 2u * hole + (left ? 1u : 2u);
            const entry w = left ? eL : eR;
            if (DEEP) set(hole, w); else arr[q] = w;
            hole = c; moved = w;
        }
    }
    // pop_heap -> __pop_heap -> __adjust_heap (stl_heap.h:224-267) followed by the caller's pop_back; all lanes, converged
    __device__ __forceinline__ void pop(unsigned lane) {
        if (n > 1u) {
            const unsigned len = n - 1u;
            const entry value = get(len);
            const unsigned limit = (len - 1u) >> 1;
            unsigned hole = 0;
            entry moved = value; bool any = false;           // the entry last moved up: it is the parent of the hole
            walk<false>(s, cached, hole, limit, moved, any);                // cached levels (local index = heap index)
            while (hole < limit && 2u * hole + 2u < cached) {               // caches of fewer than seven entries
                unsigned c = 2u * (hole + 1u);
                entry ce = s[c]; const entry le = s[c - 1u];
                if (key(ce) < key(le)) { --c; ce = le; }
                s[hole] = ce; hole = c; moved = ce; any = true;
            }
            VB200_GTL(4);
            while (hole < limit) {                           // deeper: fetch the subtree under the hole, then keep walking in shared memory
                const unsigned first = 2u * hole + 1u;
                entry r[GREEDY_PF_LEVELS + 1];
                {
                    unsigned f = first;
#pragma unroll
                    for (int l = 0; l < GREEDY_PF_LEVELS; ++l) {      // level l+1 of the subtree: 2^(l+1) entries from index f
                        const unsigned sz = 2u << l;
                        if (lane < sz && f + lane < n) r[l] = g[f + lane];
                        if (l == GREEDY_PF_LEVELS - 1 && f + lane + 32u < n) r[l + 1] = g[f + lane + 32u];
                        f = 2u * f + 1u;
                    }
                }
                __syncwarp();
                {   // scratch[q], q = 1 .. 126, is the subtree in heap layout with the hole as its (absent) root q = 0
                    unsigned f = first, off = 1u;
#pragma unroll
                    for (int l = 0; l < GREEDY_PF_LEVELS; ++l) {
                        const unsigned sz = 2u << l;
                        if (lane < sz && f + lane < n) scratch[off + lane] = r[l];
                        if (l == GREEDY_PF_LEVELS - 1 && f + lane + 32u < n) scratch[off + lane + 32u] = r[l + 1];
                        off += sz; f = 2u * f + 1u;
                    }
                }
                __syncwarp();
                walk<true>(scratch, GREEDY_PF_ENTRIES - 1u, hole, limit, moved, any);
            }
            if ((len & 1u) == 0u && hole == (len - 2u) / 2u) { const unsigned c = 2u * (hole + 1u); const entry e = get(c - 1u); set(hole, e); hole = c - 1u; moved = e; any = true; }
            if (any && !(key(moved) < key(value))) set(hole, value);      // __push_heap stops at once: the usual case
            else sift_up(hole, value);
            VB200_GTL(5);
        }
        --n;
        __syncwarp();
    }
    // ancestors of the two leaves the next push2 will add (positions n and n + 1); loads only — nothing waits for them here
    __device__ __forceinline__ void fetch_paths(unsigned lane, Paths& p) const {
        const unsigned qa = (n + 1u) >> lane, qb = (n + 2u) >> lane;      // 1-based index of the lane-th ancestor (lane 0: the leaf itself)
        p.va = lane > 0u && qa >= 1u; p.vb = lane > 0u && qb >= 1u;
        p.ia = qa - 1u; p.ib = qb - 1u;
        if (p.va) p.a = get(p.ia);
        if (p.vb) p.b = get(p.ib);
    }
    // two push_heap calls (stl_heap.h:135-148, 159-167) over the fetched paths
    __device__ __forceinline__ void push2(unsigned lane, Paths& p, entry va, entry vb) {
        const unsigned pa = n + 1u, pb = n + 2u;
        unsigned mask = __ballot_sync(0xffffffffu, lane == 0u || (p.va && key(p.a) < key(va)));
        unsigned stop = unsigned(__ffs(int(~mask))) - 1u;      // ancestors 1 .. stop-1 move down one level, the leaf's value takes ancestor stop-1's place
        if (lane >= 1u && lane < stop) set((pa >> (lane - 1u)) - 1u, p.a);
        if (lane == 0u) set((pa >> (stop - 1u)) - 1u, va);
        // the second path as the first push left it: its lane-th ancestor has the depth of the first path's ancestor j = lane - dd
        const unsigned dd = ((pb & (pb - 1u)) == 0u) ? 1u : 0u;        // the second leaf opens a new level
        const entry next = shfl1(p.a, lane + 1u - dd);                 // the first path's ancestor j + 1
        const unsigned j = lane - dd;
        if (p.vb && lane >= dd && j < stop && ((pa >> j) - 1u) == p.ib) p.b = (j == stop - 1u) ? va : next;
        mask = __ballot_sync(0xffffffffu, lane == 0u || (p.vb && key(p.b) < key(vb)));
        stop = unsigned(__ffs(int(~mask))) - 1u;
        if (lane >= 1u && lane < stop) set((pb >> (lane - 1u)) - 1u, p.b);
        if (lane == 0u) set((pb >> (stop - 1u)) - 1u, vb);
        n += 2u;
        __syncwarp();
    }
};

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

// T = float or double (the Float of the range); MIXED selects error_heuristic_mixed (double keys).  Keys are doubles whenever T is double or MIXED.
// ONEWARP: regions small enough for ONE worker warp (dimension <= 2: at most 32 evaluations per split and 32 error lines) — the CTA is two
// warps and the worker's steps are separated by __syncwarp instead of named barriers; otherwise seven worker warps.
//
// One iteration (regions-generator-adaptive-heap.h:33-41), two CTA barriers:
//   B1  the next region to split (id, dimension, where its samples are) is published, the previous children are on their way to global memory
//       heap warp:  pop (does not depend on the split: heap.front() is copied first, :33-35) -> load the paths of the two coming pushes and
//                   the samples of heap[0], the likeliest next top
//       workers:    parent -> split (the (S-1) S^(D-1) new evaluations) -> both children's ranges, errors along every dimension, heuristics
//   B2  heap warp:  two pushes over the fetched paths -> next top = heap[0]: either the region fetched above or one of the two children, whose
//                   samples are still in shared memory — the workers never wait for global memory
//       workers:    store the children
template<class F, int DIM, int SH, int SL, bool EXACT, class T = float, bool MIXED = false, bool ONEWARP = false>
__global__ void __launch_bounds__(GREEDY_THREADS, 1)
greedy_kernel(const F f, const vb200_greedy_launch a, const int heap_cached) {
    using Sh = GreedyShape<SH, SL, DIM>;
    constexpr bool KEY64 = MIXED || sizeof(T) == 8;
    using E = typename std::conditional<KEY64, GreedyEntry64, GreedyEntry32>::type;
    using Heap = GreedyHeapWarp<E>;
    using Key = typename E::key_type;
    constexpr int NPRE = (Sh::SD + 31) / 32;              // samples of the prefetched region per lane of the heap warp
    constexpr bool PREFETCH = NPRE <= 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typename E::type* s_heap = reinterpret_cast<typename E::type*>(smem_raw);
    typename E::type* s_scratch = s_heap + heap_cached;                 // [GREEDY_PF_ENTRIES]
    T* s_parent = reinterpret_cast<T*>(s_scratch + GREEDY_PF_ENTRIES);  // [SD]
    T* s_child = s_parent + Sh::SD;                                     // [2][SD]
    T* s_work = s_child + 2 * Sh::SD;                                   // [2*DIM][L]
    T* s_prange = s_work + 2 * DIM * Sh::L;                             // [2*DIM]
    T* s_crange = s_prange + 2 * DIM;                                   // [2][2*DIM]
    T* s_E = s_crange + 4 * DIM;                                        // [2][DIM]
    T* s_vol = s_E + 2 * DIM;                                           // [2]
    T* s_pre = s_vol + 2;                                               // [SD] samples of the prefetched region
    T* s_prerange = s_pre + Sh::SD;                                     // [2*DIM]
    __shared__ unsigned s_top_id, s_top_dim, s_top_src;                 // src: 0 = s_pre, 1 / 2 = child 0 / 1 of the last split, 3 = global memory
    __shared__ unsigned s_pick_dim[2];
    __shared__ double s_pick_key[2];          // heap keys of the two children (exactly representable: Key is float or double)

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, NT = blockDim.x, nwarps = NT / 32;
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

    // ---- initial region over the whole range (regions-generator-adaptive-heap.h:27-31): built as "child 0" ----
    Heap heap;                                // the heap warp's registers
    heap.g = static_cast<typename E::type*>(a.heap); heap.s = s_heap; heap.scratch = s_scratch; heap.cached = unsigned(heap_cached); heap.n = 0;
    if (tid < 2 * DIM) s_crange[tid] = tid < DIM ? rmin_[tid] : rmax_[tid - DIM];
    __syncthreads();
    for (int k = tid; k < Sh::SD; k += NT) {
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
    if (warp == 0) {
        Key err; unsigned dim; pick(s_E, s_crange, &err, &dim);           // every lane: the same values
        s_heap[0] = E::make(0u | (dim << 28), err); heap.n = 1u;
        if (lane == 0) {
            if constexpr (KEY64) static_cast<double*>(a.key64)[0] = double(err); else g_err[0] = T(err);
            s_top_id = 0u; s_top_dim = dim; s_top_src = 1u;
        }
    }
    for (int k = tid; k < Sh::SD; k += NT) g_data[k] = s_child[k];
    if (tid < 2 * DIM) g_range[tid] = s_crange[tid];

    // ---- iterations ----
    const unsigned workers = NT - 32u;
    auto worker_sync = [&] () { if constexpr (ONEWARP) __syncwarp(); else asm volatile("bar.sync 1, %0;" :: "r"(workers) : "memory"); };
    unsigned long long next_slot = 1;
#ifdef VB200_GREEDY_TIMING
    if (tid == 0) { for (int k = 0; k < 8; ++k) vb200_gt_acc[k] = 0; vb200_gt_prev = clock64(); }
#endif
    for (unsigned long long it = 0; it < a.iterations; ++it) {
        __syncthreads();                                                   // B1
        VB200_GT(0);
        const unsigned top = s_top_id, src = s_top_src; const int dim = int(s_top_dim);
        typename Heap::Paths paths;
        T pre[PREFETCH ? NPRE : 1]; T pre_range = T(0); unsigned cand = 0xffffffffu;
        if (warp == 0) {
            heap.pop(lane);
            heap.fetch_paths(lane, paths);
            if (PREFETCH && heap.n > 0u) {                                  // heap[0]: the next top unless one of the children beats it
                cand = E::id_dim(s_heap[0]) & GREEDY_ID_MASK;
#pragma unroll
                for (int q = 0; q < NPRE; ++q) if (q * 32 + int(lane) < Sh::SD) pre[q] = g_data[static_cast<unsigned long long>(cand) * Sh::SD + q * 32 + lane];
                if (lane < 2 * DIM) pre_range = g_range[static_cast<unsigned long long>(cand) * (2 * DIM) + lane];
            }
            VB200_GT(1);
        } else {
            const int wt = int(tid) - 32;
            {   // the parent's samples and range: shared memory unless the heap warp could not prefetch (src 3)
                const T* pd = src == 0u ? s_pre : src == 3u ? g_data + static_cast<unsigned long long>(top) * Sh::SD : s_child + (src - 1u) * Sh::SD;
                const T* pr = src == 0u ? s_prerange : src == 3u ? g_range + static_cast<unsigned long long>(top) * (2 * DIM) : s_crange + (src - 1u) * 2 * DIM;
                T pv = T(0), prv = T(0);
                if constexpr (ONEWARP) {                                    // s_child is both source and destination of this iteration: read, then write
                    if (wt < Sh::SD) pv = pd[wt];
                    if (wt < 2 * DIM) prv = pr[wt];
                    __syncwarp();
                    if (wt < Sh::SD) s_parent[wt] = pv;
                    if (wt < 2 * DIM) s_prange[wt] = prv;
                } else {
                    for (int k = wt; k < Sh::SD; k += int(workers)) s_parent[k] = pd[k];
                    if (wt < 2 * DIM) s_prange[wt] = pr[wt];
                }
            }
            worker_sync();
            // split along `dim` (split.h:13-49): the (2S-1)-wide array; even positions are the parent's samples, odd ones new evaluations
            int inner = 1; for (int i = 0; i < dim; ++i) inner *= SH;
            constexpr int NEW = (SH - 1) * Sh::L;
            for (int idx = wt; idx < NEW; idx += int(workers)) {            // the new evaluations first: they are what takes time
                const int i = 2 * (idx / Sh::L) + 1, o = idx % Sh::L;      // position along `dim` (odd), index over the other dims
                const int lo = o % inner, hi = o / inner;
                std::array<T, DIM> x; int t = o;
#pragma unroll
                for (int d = 0; d < DIM; ++d) {
                    double p;
                    if (d == dim) p = frac(i, 2 * (SH - 1));
                    else { p = frac(t % SH, SH - 1); t /= SH; }
                    x[d] = grid_coord<T>(p, s_prange[d], s_prange[DIM + d]);
                }
                const T v = f(x);
                if (i < SH - 1) s_child[lo + i * inner + hi * inner * SH] = v;
                else s_child[Sh::SD + lo + (i - (SH - 1)) * inner + hi * inner * SH] = v;
            }
            for (int idx = wt; idx < Sh::SD; idx += int(workers)) {         // the parent's own samples: even positions 2e
                const int e = idx / Sh::L, o = idx % Sh::L, i = 2 * e;
                const int lo = o % inner, hi = o / inner;
                const T v = s_parent[lo + e * inner + hi * inner * SH];
                if (i <= SH - 1) s_child[lo + i * inner + hi * inner * SH] = v;
                if (i >= SH - 1) s_child[Sh::SD + lo + (i - (SH - 1)) * inner + hi * inner * SH] = v;
            }
            // child ranges (region.h:349-357): d = (max-min)/Float(2); child 0 = [min, min+d*1], child 1 = [min+d*1, max]
            if (wt >= int(workers) - 2) {
                const int c = wt - (int(workers) - 2);
                const T pmin = s_prange[dim], pmax = s_prange[DIM + dim];
                const T mid = rules::add(pmin, rules::mul(rules::quo(rules::sub(pmax, pmin), T(2)), T(1)));
                T* cr = s_crange + c * 2 * DIM;
                for (int d = 0; d < 2 * DIM; ++d) cr[d] = s_prange[d];
                if (c == 0) cr[DIM + dim] = mid; else cr[dim] = mid;
                T v = T(1); for (int d = 0; d < DIM; ++d) v = rules::mul(v, rules::sub(cr[DIM + d], cr[d]));
                s_vol[c] = v;
            }
            worker_sync();
            // nested-rule error of both children along every dimension (region.h:387-393)
            if constexpr (ONEWARP) {
                // one lane per (child, dimension, line): L lines per job, 2*DIM*L <= 32; the fold over the other dimension is one rule application
                const int job = wt / Sh::L, o = wt % Sh::L;
                if (job < 2 * DIM) {
                    const int c = job / DIM, d = job % DIM;
                    int in_d = 1; for (int i = 0; i < d; ++i) in_d *= SH;
                    const int lo = o % in_d, hi = o / in_d;
                    T line[SH];
#pragma unroll
                    for (int e = 0; e < SH; ++e) line[e] = s_child[c * Sh::SD + lo + e * in_d + hi * in_d * SH];
                    s_work[job * Sh::L + o] = rules::line_error<SH, SL, T>(rel_of(d), line);
                }
                __syncwarp();
                if (wt < 2 * DIM) {
                    const T* w = s_work + wt * Sh::L;
                    const T folded = Sh::L == 1 ? w[0] : rules::apply<SH, T>(w);
                    s_E[wt] = rules::mul(s_vol[wt / DIM], folded);
                }
            } else {
                for (int job = int(warp) - 1; job < 2 * DIM; job += int(nwarps) - 1) {      // one worker warp per (child, dimension)
                    const int c = job / DIM, d = job % DIM;
                    const T e = region_error_warp<SH, SL, DIM, T>(s_child + c * Sh::SD, s_vol[c], d, rel_of(d), s_work + job * Sh::L, lane);
                    if (lane == 0) s_E[c * DIM + d] = e;
                }
            }
            worker_sync();
            if (wt < 2) {                                                   // the two children's heuristics side by side (:36-40)
                Key err; unsigned d; pick(s_E + wt * DIM, s_crange + wt * 2 * DIM, &err, &d);
                s_pick_key[wt] = double(err); s_pick_dim[wt] = d;
            }
        }
        __syncthreads();                                                   // B2: popped; children, errors and picks ready
        VB200_GT(2);
        if (warp == 0) {
            const typename E::type va = E::make(unsigned(next_slot) | (s_pick_dim[0] << 28), Key(s_pick_key[0]));
            const typename E::type vb = E::make(unsigned(next_slot + 1) | (s_pick_dim[1] << 28), Key(s_pick_key[1]));
            heap.push2(lane, paths, va, vb);
            const unsigned idd = E::id_dim(s_heap[0]), id = idd & GREEDY_ID_MASK;
            unsigned nsrc = 3u;
            if (id == unsigned(next_slot)) nsrc = 1u; else if (id == unsigned(next_slot + 1)) nsrc = 2u;
            else if (PREFETCH && id == cand) {
                nsrc = 0u;
#pragma unroll
                for (int q = 0; q < NPRE; ++q) if (q * 32 + int(lane) < Sh::SD) s_pre[q * 32 + lane] = pre[q];
                if (lane < 2 * DIM) s_prerange[lane] = pre_range;
            }
            if (lane == 0) { s_top_id = id; s_top_dim = idd >> 28; s_top_src = nsrc; }
            VB200_GT(3);
        } else {
            // store the children (slots next_slot, next_slot+1)
            const int wt = int(tid) - 32;
            for (int k = wt; k < 2 * Sh::SD; k += int(workers)) g_data[next_slot * Sh::SD + k] = s_child[k];
            if (wt < 4 * DIM) g_range[next_slot * (2 * DIM) + wt] = s_crange[wt];
            if (wt < 2) { if constexpr (KEY64) static_cast<double*>(a.key64)[next_slot + wt] = s_pick_key[wt]; else g_err[next_slot + wt] = T(s_pick_key[wt]); }
        }
        next_slot += 2;
    }
    __syncthreads();
#ifdef VB200_GREEDY_TIMING
    if (tid == 0) for (int k = 0; k < 8; ++k) vb200_greedy_clock[k] = vb200_gt_acc[k];
#endif
    // flush the cached top of the heap
    __shared__ unsigned s_n;
    if (tid == 0) { *a.heap_size = static_cast<uint64_t>(heap.n); s_n = heap.n; }
    __syncthreads();
    for (unsigned i = tid; i < s_n && i < unsigned(heap_cached); i += NT) static_cast<typename E::type*>(a.heap)[i] = s_heap[i];
}

template<class F, int DIM, int SH, int SL, bool EXACT, class T, bool MIXED>
inline int launch_greedy_rule(const F& f, const vb200_greedy_launch& a, cudaStream_t st) {
    using Sh = GreedyShape<SH, SL, DIM>;
    constexpr bool KEY64 = MIXED || sizeof(T) == 8;
    constexpr size_t ENTRY = KEY64 ? 16 : 8;
    // one worker warp: at most 32 new evaluations per split, 32 samples per region and 32 error lines, fold = one rule application
    constexpr bool ONEWARP = DIM <= 2 && Sh::SD <= 32 && (SH - 1) * Sh::L <= 32 && 2 * DIM * Sh::L <= 32;
    if (a.capacity > GREEDY_ID_MASK) return int(cudaErrorInvalidValue);
    auto k = greedy_kernel<F, DIM, SH, SL, EXACT, T, MIXED, ONEWARP>;
    const size_t fixed = sizeof(T) * size_t(4 * Sh::SD + 2 * DIM * Sh::L + 2 * DIM + 4 * DIM + 2 * DIM + 2 + 2 * DIM) + size_t(GREEDY_PF_ENTRIES) * ENTRY + 64;
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
    k<<<1, ONEWARP ? 64 : GREEDY_THREADS, smem, st>>>(f, a, cached);
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
