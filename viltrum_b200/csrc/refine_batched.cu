// Batched adaptive refinement — the throughput mode of the region generator (SURVEY.md §2.2 K5-K7 "batched", §7 step 6).
// The reference refines greedily, one region per iteration (regions-generator-adaptive-heap.h:32-42); that loop is serial by
// construction (greedy.cuh reproduces it exactly).  Here every round
//   1. selects the B regions with the largest error heuristic with a device-side radix top-k over the float keys
//      (four 8-bit histogram passes find the B-th largest key, an ordered compaction takes everything above it plus the
//      first ties in table order),
//   2. generates the (S-1)*S^(D-1) new sample points of all B splits at once, evaluates the integrand on them through the
//      eval thunk (one launch), and
//   3. builds both children of every split and their error heuristics, one warp per split, with the same arithmetic as the
//      exact mode (rules.cuh) — a region produced here is bit-identical to the one the greedy mode would produce for the
//      same split.
// Round sizes follow from the host-known region count (B = max(1, n/4) unless params.batch caps it), so the whole
// refinement is enqueued without a single device->host read.  With batch > 1 the leaf SET differs from the reference's
// (equally valid: > 99 % of the keys are ties at C3 scale), so this mode is validated by convergence,
// not bin by bin.
#include "regions.h"
#include <viltrum_b200/device/greedy.cuh>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cub/device/device_radix_sort.cuh>

using namespace vb200;
namespace R = viltrum::b200::device::rules;
namespace D_ = viltrum::b200::device;

namespace {

struct SelectState { unsigned prefix_val, prefix_mask; unsigned long long k; };

// Radix top-k over the float keys' bits, eight bits per pass, most significant first.  Every pass has its own 256-bin histogram (hist[pass][256]),
// so no kernel is needed between two passes: a CTA that starts pass p replays the p digit picks made so far from the finished histograms
// (a 256-thread suffix scan each) instead of reading a state a single-CTA launch would have had to write; the count kernel replays all four and
// its first CTA writes the final state for the write kernel.
__global__ void select_init_kernel(SelectState* st, unsigned long long k, unsigned* hist) {
    if (threadIdx.x == 0) { st->prefix_val = 0; st->prefix_mask = 0; st->k = k; }
    for (unsigned i = threadIdx.x; i < 4 * 256; i += blockDim.x) hist[i] = 0;
}
// The digit d whose bucket holds the k-th largest key among the keys of one histogram: the largest d >= 1 with hist[d] + ... + hist[255] >= k,
// else 0 (what a serial walk from 255 down finds); k is reduced by the keys above that digit.  All 256 threads of the CTA call it.
__device__ __forceinline__ void select_pick_digit(const unsigned* __restrict__ hist, int shift, unsigned long long& k, unsigned& pv, unsigned& pm,
                                                  unsigned long long* s_s, int* s_d, unsigned long long* s_cum) {
    const unsigned d = threadIdx.x;
    const unsigned h = hist[d];
    __syncthreads();                                       // the previous pick's readers are done with the scratch
    s_s[d] = h;
    if (d == 0) *s_d = 0;
    __syncthreads();
    for (unsigned off = 1; off < 256; off <<= 1) {         // inclusive suffix sums (Hillis-Steele)
        const unsigned long long v = d + off < 256 ? s_s[d + off] : 0ull;
        __syncthreads();
        s_s[d] += v;
        __syncthreads();
    }
    if (d >= 1 && s_s[d] >= k) atomicMax(s_d, int(d));
    __syncthreads();
    if (d == unsigned(*s_d)) *s_cum = s_s[d] - h;          // keys in the digits above d
    __syncthreads();
    k -= *s_cum; pv |= unsigned(*s_d) << shift; pm |= 255u << shift;
}
// histogram of pass `pass` (digit at bit 24 - 8 pass) among the keys that match the prefix found by the earlier passes
__global__ void __launch_bounds__(256) select_hist_kernel(const float* __restrict__ keys, uint64_t n, unsigned long long k0, int pass, unsigned* __restrict__ hist) {
    __shared__ unsigned s_hist[256];
    __shared__ unsigned long long s_s[256], s_cum;
    __shared__ int s_d;
    s_hist[threadIdx.x] = 0;
    unsigned long long k = k0; unsigned pv = 0, pm = 0;
    for (int q = 0; q < pass; ++q) select_pick_digit(hist + q * 256, 24 - 8 * q, k, pv, pm, s_s, &s_d, &s_cum);
    __syncthreads();
    const int shift = 24 - 8 * pass;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
        const unsigned key = __float_as_uint(keys[i]);
        if ((key & pm) == pv) atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (s_hist[threadIdx.x]) atomicAdd(&hist[pass * 256 + threadIdx.x], s_hist[threadIdx.x]);
}
// per-CTA counts of keys above the threshold and equal to it (CTA = 1024 consecutive regions)
__global__ void __launch_bounds__(256) select_count_kernel(const float* __restrict__ keys, uint64_t n, unsigned long long k0, const unsigned* __restrict__ hist, SelectState* __restrict__ st,
                                                           unsigned* __restrict__ cta_gt, unsigned* __restrict__ cta_eq) {
    __shared__ unsigned s_gt, s_eq;
    __shared__ unsigned long long s_s[256], s_cum;
    __shared__ int s_d;
    if (threadIdx.x == 0) { s_gt = 0; s_eq = 0; }
    // the four digit picks once more: the threshold is complete now; CTA 0 writes it down for the write kernel
    unsigned long long k = k0; unsigned pv = 0, pm = 0;
    for (int q = 0; q < 4; ++q) select_pick_digit(hist + q * 256, 24 - 8 * q, k, pv, pm, s_s, &s_d, &s_cum);
    if (blockIdx.x == 0 && threadIdx.x == 0) { st->k = k; st->prefix_val = pv; st->prefix_mask = pm; }
    __syncthreads();
    const unsigned T = pv;
    unsigned gt = 0, eq = 0;
    for (int j = 0; j < 4; ++j) {
        const uint64_t i = uint64_t(blockIdx.x) * 1024 + j * 256 + threadIdx.x;
        if (i < n) { const unsigned kk = __float_as_uint(keys[i]); gt += kk > T; eq += kk == T; }
    }
    atomicAdd(&s_gt, gt); atomicAdd(&s_eq, eq);
    __syncthreads();
    if (threadIdx.x == 0) { cta_gt[blockIdx.x] = s_gt; cta_eq[blockIdx.x] = s_eq; }
}
// exclusive scan of the per-CTA counts (single CTA; at most a few thousand entries)
__global__ void __launch_bounds__(1024) select_scan_kernel(unsigned* cta_gt, unsigned* cta_eq, unsigned nctas) {
    __shared__ unsigned s_a[1024], s_b[1024];
    unsigned run_a = 0, run_b = 0;
    for (unsigned base = 0; base < nctas; base += 1024) {
        const unsigned i = base + threadIdx.x;
        const unsigned a = i < nctas ? cta_gt[i] : 0, b = i < nctas ? cta_eq[i] : 0;
        s_a[threadIdx.x] = a; s_b[threadIdx.x] = b;
        __syncthreads();
        for (unsigned off = 1; off < 1024; off <<= 1) {
            const unsigned va = threadIdx.x >= off ? s_a[threadIdx.x - off] : 0, vb = threadIdx.x >= off ? s_b[threadIdx.x - off] : 0;
            __syncthreads();
            s_a[threadIdx.x] += va; s_b[threadIdx.x] += vb;
            __syncthreads();
        }
        if (i < nctas) { cta_gt[i] = run_a + s_a[threadIdx.x] - a; cta_eq[i] = run_b + s_b[threadIdx.x] - b; }
        run_a += s_a[1023]; run_b += s_b[1023];
        __syncthreads();
    }
}
// ordered compaction: selected = key > T, or key == T and fewer than k_eq equal keys precede it in table order.
// Output position: (#selected-by-'>' before) + (#selected ties before); both are monotone in the index, so sel[] is sorted.
// SCANNED: cta_gt / cta_eq hold exclusive prefixes (select_scan_kernel ran); otherwise they hold the raw per-CTA counts and every CTA sums the
// entries before its own (tables of up to a few million regions: a few thousand entries — cheaper than one more launch).
template<bool SCANNED>
__global__ void __launch_bounds__(1024) select_write_kernel(const float* __restrict__ keys, uint64_t n, const SelectState* __restrict__ st,
                                                            const unsigned* __restrict__ cta_gt, const unsigned* __restrict__ cta_eq, unsigned* __restrict__ sel, unsigned* __restrict__ hist) {
    __shared__ unsigned s_wgt[32], s_weq[32], s_pg[32], s_pe[32];
    const unsigned T = st->prefix_val; const unsigned long long k_eq = st->k;
    const uint64_t i = uint64_t(blockIdx.x) * 1024 + threadIdx.x;
    const unsigned key = i < n ? __float_as_uint(keys[i]) : 0u;
    const bool gt = i < n && key > T, eq = i < n && key == T;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned mg = __ballot_sync(0xffffffffu, gt), me = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) { s_wgt[warp] = __popc(mg); s_weq[warp] = __popc(me); }
    unsigned bg = 0, be = 0;
    if constexpr (!SCANNED) {
        unsigned pg = 0, pe = 0;
        for (unsigned c = threadIdx.x; c < blockIdx.x; c += 1024) { pg += cta_gt[c]; pe += cta_eq[c]; }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) { pg += __shfl_xor_sync(0xffffffffu, pg, off); pe += __shfl_xor_sync(0xffffffffu, pe, off); }
        if (lane == 0) { s_pg[warp] = pg; s_pe[warp] = pe; }
    }
    __syncthreads();
    if constexpr (SCANNED) { bg = cta_gt[blockIdx.x]; be = cta_eq[blockIdx.x]; }
    else { for (unsigned w = 0; w < 32; ++w) { bg += s_pg[w]; be += s_pe[w]; } }
    for (unsigned w = 0; w < warp; ++w) { bg += s_wgt[w]; be += s_weq[w]; }
    bg += __popc(mg & ((1u << lane) - 1u)); be += __popc(me & ((1u << lane) - 1u));
    // ties are taken in table order: tie number `be` is selected iff be < k_eq; ties before it that were selected: min(be, k_eq)
    const unsigned long long ties_before = be < k_eq ? be : k_eq;
    if (gt) sel[bg + ties_before] = unsigned(i);
    else if (eq && be < k_eq) sel[bg + be] = unsigned(i);
    if (blockIdx.x == 0) hist[threadIdx.x] = 0;            // 4 x 256 bins: nobody reads them after the count kernel; clean for the next round
}

// The same selection for small tables in ONE launch of one CTA (keys in shared memory): the first ~35 of the ~60 rounds of a 10^6-split
// refinement hold fewer than 8192 regions and were eleven tiny launches each — launch latency, not work.  Same threshold, same ties in
// table order, hence the same sel[] as select_hist/pick/count/scan/write.
constexpr unsigned SELECT_SMALL_MAX = 8192;
__global__ void __launch_bounds__(1024) select_small_kernel(const float* __restrict__ keys, unsigned n, unsigned long long k, SelectState* st, unsigned* __restrict__ sel) {
    __shared__ unsigned s_key[SELECT_SMALL_MAX];
    __shared__ unsigned s_hist[256];
    __shared__ unsigned s_pv, s_pm, s_wgt[32], s_weq[32], s_base_gt, s_base_eq;
    __shared__ unsigned long long s_k;
    const unsigned tid = threadIdx.x;
    for (unsigned i = tid; i < n; i += 1024) s_key[i] = __float_as_uint(keys[i]);
    if (tid == 0) { s_pv = 0; s_pm = 0; s_k = k; }
    __syncthreads();
    for (int shift = 24; shift >= 0; shift -= 8) {
        if (tid < 256) s_hist[tid] = 0;
        __syncthreads();
        const unsigned pv = s_pv, pm = s_pm;
        for (unsigned i = tid; i < n; i += 1024) { const unsigned kk = s_key[i]; if ((kk & pm) == pv) atomicAdd(&s_hist[(kk >> shift) & 255u], 1u); }
        __syncthreads();
        if (tid < 32) {
            // the digit d whose bucket holds the k-th largest key: the largest d >= 1 with suffix(d) = sum_{i >= d} hist[i] >= k, else 0; cum = suffix(d+1).
            // Lane l owns buckets 8l..8l+7; a shuffle suffix scan over the lanes, then eight buckets per lane (was a 256-step loop on one thread).
            unsigned h[8]; unsigned own = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { h[q] = s_hist[8 * tid + q]; own += h[q]; }
            unsigned above = own;                                   // inclusive suffix over lanes >= tid
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const unsigned v = __shfl_down_sync(0xffffffffu, above, off); if (tid + off < 32) above += v; }
            above -= own;                                           // buckets of the lanes above this one
            const unsigned long long kk = s_k;
            int best = -1; unsigned long long best_cum = 0; unsigned long long run = above;
#pragma unroll
            for (int q = 7; q >= 0; --q) { if (best < 0 && run + h[q] >= kk) { best = 8 * int(tid) + q; best_cum = run; } run += h[q]; }
            if (best == 0) best = -1;                               // d = 0 is the fall-through, never a match (the loop stops at d > 0)
            const unsigned hit = __ballot_sync(0xffffffffu, best >= 0);
            int d = 0; unsigned long long cum = 0;
            if (hit) {
                const int src = 31 - __clz(int(hit));               // the highest lane with a match holds the largest d
                d = __shfl_sync(0xffffffffu, best, src); cum = __shfl_sync(0xffffffffu, best_cum, src);
            } else {                                                // everything above bucket 0
                cum = __shfl_sync(0xffffffffu, run, 0) - s_hist[0];
            }
            __syncwarp();                                           // every lane has read s_k (racecheck: the shuffles above order execution, not memory)
            if (tid == 0) { s_k = kk - cum; s_pv |= unsigned(d) << shift; s_pm |= 255u << shift; }
        }
        __syncthreads();
    }
    const unsigned T = s_pv; const unsigned long long k_eq = s_k;
    if (tid == 0) { s_base_gt = 0; s_base_eq = 0; st->prefix_val = T; st->prefix_mask = s_pm; st->k = k_eq; }
    __syncthreads();
    const unsigned lane = tid & 31u, warp = tid >> 5;
    for (unsigned base = 0; base < n; base += 1024) {
        const unsigned i = base + tid;
        const unsigned key = i < n ? s_key[i] : 0u;
        const bool gt = i < n && key > T, eq = i < n && key == T;
        const unsigned mg = __ballot_sync(0xffffffffu, gt), me = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) { s_wgt[warp] = __popc(mg); s_weq[warp] = __popc(me); }
        __syncthreads();
        unsigned bg = s_base_gt, be = s_base_eq;
        for (unsigned w = 0; w < warp; ++w) { bg += s_wgt[w]; be += s_weq[w]; }
        bg += __popc(mg & ((1u << lane) - 1u)); be += __popc(me & ((1u << lane) - 1u));
        const unsigned long long ties_before = be < k_eq ? be : k_eq;
        if (gt) sel[bg + ties_before] = i;
        else if (eq && be < k_eq) sel[bg + be] = i;
        __syncthreads();
        if (tid == 1023) { s_base_gt = bg + (gt ? 1u : 0u); s_base_eq = be + (eq ? 1u : 0u); }
        __syncthreads();
    }
}

// sample points of all selected splits: point q of split r = odd position i = 2*(q / L)+1 along the split dimension,
// other-dims index o = q % L; coordinates from the PARENT range (split.h:22, region.h:40-46)
// i / m for the grid positions (region.h:40-46): m = S-1 or 2(S-1) is a power of two for S = 3, 5, so the double division is an exact multiplication
__device__ __forceinline__ double grid_frac(int i, int m) { return ((m & (m - 1)) == 0) ? R::dm(double(i), 1.0 / double(m)) : R::dd(double(i), double(m)); }
template<int S, int DIM>
__global__ void split_points_kernel(uint64_t cap, uint64_t nsel, const unsigned* __restrict__ sel,
                                    const float* __restrict__ rmin, const float* __restrict__ rmax, const uint32_t* __restrict__ errdim, float* __restrict__ points) {
    constexpr int L = R::ipow(S, DIM - 1);
    constexpr unsigned Q = unsigned(S - 1) * L;
    const uint64_t N = nsel * Q;
    const uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= N) return;
    const uint64_t r = N <= 0xffffffffull ? uint64_t(unsigned(t) / Q) : t / Q;         // Q is a compile-time constant: multiply-shift, 32-bit where it fits
    const int q = int(t - r * Q);
    const unsigned slot = sel[r]; const int sd = int(errdim[slot]);
    const int i = 2 * (q / L) + 1; int o = q % L;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        double p;
        if (d == sd) p = grid_frac(i, 2 * (S - 1));
        else { p = grid_frac(o % S, S - 1); o /= S; }
        points[uint64_t(d) * N + t] = D_::grid_coord(p, rmin[uint64_t(d) * cap + slot], rmax[uint64_t(d) * cap + slot]);
    }
}

// heuristic of a region from its per-dimension errors: error_heuristic_default / _size (float keys) or _mixed (a double key upstream, rounded
// to float here: the batched mode orders by float keys and is not bit-exact with the greedy order anyway)
struct HeurArgs { int heuristic, metric, metric_rest; double size_weight; D_::MixedParams mixed; };
inline HeurArgs heur_args(int heuristic, int metric, double size_weight, const vb200_mixed_heuristic& m) {
    HeurArgs h; h.heuristic = heuristic; h.metric = metric; h.metric_rest = m.metric_rest; h.size_weight = size_weight;
    h.mixed = D_::MixedParams{m.dimension, m.bins_weight, size_weight, m.size_threshold_bins, m.size_threshold_rest, m.error_increase_factor};
    return h;
}
__device__ __forceinline__ bool heur_relative(const HeurArgs& h, int d) {
    if (h.heuristic == VB200_HEURISTIC_MIXED) return ((d == 0 || d < h.mixed.dimension) ? h.metric : h.metric_rest) == VB200_METRIC_RELATIVE;
    return h.metric == VB200_METRIC_RELATIVE;
}
template<int DIM>
__device__ __forceinline__ void heur_pick(const HeurArgs& h, const float* E, const float* rng, float* e, unsigned* d) {
    if (h.heuristic == VB200_HEURISTIC_MIXED) { double k; D_::heuristic_pick_mixed<DIM, float>(E, rng, h.mixed, &k, d); *e = float(k); }
    else D_::heuristic_pick<DIM>(E, rng, h.heuristic, h.size_weight, e, d);
}

// one GROUP of threads per split — a warp (CTA = false: the throughput form, 8 splits per CTA) or a whole CTA (CTA = true: rounds with few
// splits, where a lone warp's ~30 us of dependent work per split is the round's whole duration): build both children (region.h:345-359,
// split.h:13-49), store child 0 over the parent and child 1 in a new slot, then evaluate both children's error heuristics
// (region.h:387-411, error-heuristic.h:10-46).  Same operations on the same operands either way, hence the same bits.
template<int SH, int SL, int DIM, bool CTA>
__global__ void split_children_kernel(uint64_t cap, uint64_t nsel, uint64_t n_old, const unsigned* __restrict__ sel, const float* __restrict__ vals,
                                      float* __restrict__ rmin, float* __restrict__ rmax, float* __restrict__ data, float* __restrict__ err, uint32_t* __restrict__ errdim,
                                      const HeurArgs hr) {
    using Sh = D_::GreedyShape<SH, SL, DIM>;
    extern __shared__ float smem[];
    const unsigned warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const unsigned lane = CTA ? threadIdx.x : (threadIdx.x & 31u);          // index within the group
    const int G = CTA ? int(blockDim.x) : 32;                                // threads of the group
    auto gsync = [] () { if constexpr (CTA) __syncthreads(); else __syncwarp(); };
    float* s_parent = smem + (CTA ? size_t(0) : size_t(warp) * (3 * Sh::SD + 2 * DIM * Sh::L + 4 * DIM + 2 * DIM + 2));
    float* s_child = s_parent + Sh::SD;
    float* s_work = s_child + 2 * Sh::SD;      // [2*DIM][L]
    float* s_crange = s_work + 2 * DIM * Sh::L;          // [2][2*DIM]
    float* s_E = s_crange + 4 * DIM;           // [2][DIM]
    float* s_vol = s_E + 2 * DIM;              // [2]
    const uint64_t r = CTA ? uint64_t(blockIdx.x) : uint64_t(blockIdx.x) * wpc + warp;
    if (r >= nsel) return;
    const unsigned slot = sel[r]; const uint64_t slot1 = n_old + r;
    const int dim = int(errdim[slot]);
    for (int k = lane; k < Sh::SD; k += G) s_parent[k] = data[uint64_t(k) * cap + slot];
    float prange[2 * DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) { prange[d] = rmin[uint64_t(d) * cap + slot]; prange[DIM + d] = rmax[uint64_t(d) * cap + slot]; }
    gsync();
    int inner = 1; for (int i = 0; i < dim; ++i) inner *= SH;
    constexpr int Q = (SH - 1) * Sh::L;
    for (int item = lane; item < Sh::WIDE; item += G) {
        const int i = item / Sh::L, o = item % Sh::L;
        const int lo = o % inner, hi = o / inner;
        const float v = (i & 1) == 0 ? s_parent[lo + (i / 2) * inner + hi * inner * SH] : vals[r * uint64_t(Q) + uint64_t((i / 2) * Sh::L + o)];
        if (i <= SH - 1) s_child[lo + i * inner + hi * inner * SH] = v;
        if (i >= SH - 1) s_child[Sh::SD + lo + (i - (SH - 1)) * inner + hi * inner * SH] = v;
    }
    if (lane < 2) {
        const float pmin = prange[dim], pmax = prange[DIM + dim];
        const float mid = R::fa(pmin, R::fm(R::fd(R::fs(pmax, pmin), 2.0f), 1.0f));
        float* cr = s_crange + lane * 2 * DIM;
        for (int d = 0; d < 2 * DIM; ++d) cr[d] = prange[d];
        if (lane == 0) cr[DIM + dim] = mid; else cr[dim] = mid;
        float v = 1.0f; for (int d = 0; d < DIM; ++d) v = R::fm(v, R::fs(cr[DIM + d], cr[d]));
        s_vol[lane] = v;
    }
    gsync();
    for (int k = lane; k < Sh::SD; k += G) { data[uint64_t(k) * cap + slot] = s_child[k]; data[uint64_t(k) * cap + slot1] = s_child[Sh::SD + k]; }
    if (lane < DIM) {
        rmin[uint64_t(lane) * cap + slot] = s_crange[lane]; rmax[uint64_t(lane) * cap + slot] = s_crange[DIM + lane];
        rmin[uint64_t(lane) * cap + slot1] = s_crange[2 * DIM + lane]; rmax[uint64_t(lane) * cap + slot1] = s_crange[3 * DIM + lane];
    }
    // nested-rule error of both children along every dimension (region.h:387-393), all 2*DIM of them side by side: the line errors are
    // 2*DIM*L independent tasks, every level of the fold_all that follows 2*DIM*n — instead of ten folds one after the other, most of whose
    // steps keep a handful of lanes busy.  Every value is computed by the same operations as in region_error_warp, hence the same bits.
    constexpr int CD = 2 * DIM;
    for (int t = lane; t < CD * Sh::L; t += G) {
        const int cd = t / Sh::L, o = t % Sh::L, c = cd / DIM, d = cd % DIM;
        int inn = 1; for (int i = 0; i < d; ++i) inn *= SH;
        const int lo = o % inn, hi = o / inn;
        float line[SH];
#pragma unroll
        for (int e = 0; e < SH; ++e) line[e] = s_child[c * Sh::SD + lo + e * inn + hi * inn * SH];
        s_work[cd * Sh::L + o] = R::line_error<SH, SL>(heur_relative(hr, d), line);
    }
    gsync();
    for (int n = Sh::L / SH; n >= 1; n /= SH) {                     // fold_all(high rule): fold dimension 0 until one value is left (fold.h:87-108)
        constexpr int MAXT = (CD * (Sh::L / SH) + 31) / 32;
        float v[MAXT > 0 ? MAXT : 1];
        int k = 0;
        for (int t = lane; t < CD * n; t += G, ++k) { const int cd = t / n, o = t % n; v[k] = R::apply<SH>(s_work + cd * Sh::L + o * SH); }
        gsync();
        k = 0;
        for (int t = lane; t < CD * n; t += G, ++k) { const int cd = t / n, o = t % n; s_work[cd * Sh::L + o] = v[k]; }
        gsync();
        if (n == 1) break;
    }
    if (lane < CD) s_E[lane] = R::fm(s_vol[lane / DIM], s_work[lane * Sh::L]);
    gsync();
    if (lane < 2) {
        float e; unsigned d;
        heur_pick<DIM>(hr, s_E + lane * DIM, s_crange + lane * 2 * DIM, &e, &d);
        const uint64_t s = lane == 0 ? uint64_t(slot) : slot1;
        err[s] = e; errdim[s] = d;
    }
}
// rounds of at most this many splits give every split a CTA (VB200_SPLIT_CTA_MAX overrides: tuning / test knob)
constexpr uint64_t SPLIT_CTA_MAX = 4736;      // measured on C4 (ms per step): 0: 10.76, 296: 10.18, 1184: 10.05, 4736: 10.00, always: 10.01
inline uint64_t split_cta_max() {
    if (const char* e = std::getenv("VB200_SPLIT_CTA_MAX")) return std::strtoull(e, nullptr, 10);
    return SPLIT_CTA_MAX;
}

// the root region: samples are already in data[.][0]; compute its heuristic
template<int SH, int SL, int DIM>
__global__ void root_error_kernel(uint64_t cap, const float* __restrict__ rmin, const float* __restrict__ rmax, const float* __restrict__ data,
                                  float* __restrict__ err, uint32_t* __restrict__ errdim, const HeurArgs hr) {
    using Sh = D_::GreedyShape<SH, SL, DIM>;
    extern __shared__ float smem[];
    float* s_data = smem; float* s_work = s_data + Sh::SD; float* s_range = s_work + Sh::L; float* s_E = s_range + 2 * DIM;
    const unsigned lane = threadIdx.x;
    for (int k = lane; k < Sh::SD; k += 32) s_data[k] = data[uint64_t(k) * cap];
    if (lane < DIM) { s_range[lane] = rmin[uint64_t(lane) * cap]; s_range[DIM + lane] = rmax[uint64_t(lane) * cap]; }
    __syncwarp();
    float vol = 1.0f; for (int d = 0; d < DIM; ++d) vol = R::fm(vol, R::fs(s_range[DIM + d], s_range[d]));
    for (int d = 0; d < DIM; ++d) {
        const float e = D_::region_error_warp<SH, SL, DIM>(s_data, vol, d, heur_relative(hr, d), s_work, lane);
        if (lane == 0) s_E[d] = e;
        __syncwarp();
    }
    if (lane == 0) { float e; unsigned d; heur_pick<DIM>(hr, s_E, s_range, &e, &d); err[0] = e; errdim[0] = d; }
}

template<int SH, int SL, int DIM>
int run_rounds(vb200_ctx* ctx, const vb200_integrand* f, const vb200_adaptive_params* p, vb200_regions* r,
               SelectState* st, unsigned* hist, unsigned* cta_gt, unsigned* cta_eq, unsigned* sel, float* points, float* vals, uint64_t max_batch) {
    using Sh = D_::GreedyShape<SH, SL, DIM>;
    const uint64_t cap = r->capacity;
    cudaStream_t s = ctx->stream;
    root_error_kernel<SH, SL, DIM><<<1, 32, (Sh::SD + Sh::L + 4 * DIM) * sizeof(float), s>>>(cap, r->rmin, r->rmax, r->data, r->err, r->errdim, heur_args(p->heuristic,
            p->metric, p->size_weight, p->mixed));
    ctx->launches++;
    VB200_CUDA(ctx, cudaGetLastError());
    const size_t per_warp = (3 * Sh::SD + 2 * DIM * Sh::L + 4 * DIM + 2 * DIM + 2) * sizeof(float);
    int wpc = int((96u << 10) / per_warp); if (wpc > 8) wpc = 8; if (wpc < 1) wpc = 1;
    auto kchild = split_children_kernel<SH, SL, DIM, false>;
    auto kchild_cta = split_children_kernel<SH, SL, DIM, true>;
    VB200_CUDA(ctx, cudaFuncSetAttribute(kchild, cudaFuncAttributeMaxDynamicSharedMemorySize, int(per_warp * wpc)));
    VB200_CUDA(ctx, cudaFuncSetAttribute(kchild_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, int(per_warp)));
    const uint64_t cta_max = split_cta_max();
    const uint64_t Q = uint64_t(SH - 1) * Sh::L;
    uint64_t n = 1, left = p->iterations;
    uint64_t small_max = SELECT_SMALL_MAX;
    if (const char* env = std::getenv("VB200_SELECT_SMALL_MAX")) small_max = std::min<uint64_t>(std::strtoull(env, nullptr, 10), SELECT_SMALL_MAX);     // test knob
    bool hist_clean = false;
    unsigned fused_max = 4096;      // per-CTA counts up to which the write kernel sums its own prefix (VB200_SELECT_FUSED_MAX: test knob for the scan kernel's path)
    if (const char* env = std::getenv("VB200_SELECT_FUSED_MAX")) fused_max = unsigned(std::strtoul(env, nullptr, 10));
    while (left > 0) {
        uint64_t B = n / 4; if (B < 1) B = 1; if (B > left) B = left; if (B > max_batch) B = max_batch;
        if (p->batch > 1 && B > uint64_t(p->batch)) B = uint64_t(p->batch);
        // 1. radix top-k
        if (n <= small_max) {
            select_small_kernel<<<1, 1024, 0, s>>>(r->err, unsigned(n), B, st, sel);
            ctx->launches += 1;
        } else {
            if (!hist_clean) { select_init_kernel<<<1, 256, 0, s>>>(st, B, hist); ctx->launches++; hist_clean = true; }      // later rounds: the write kernel leaves the histograms zeroed
            const unsigned hgrid = unsigned(std::min<uint64_t>((n + 255) / 256, uint64_t(ctx->sm_count) * 8));
            for (int pass = 0; pass < 4; ++pass) select_hist_kernel<<<hgrid, 256, 0, s>>>(r->err, n, B, pass, hist);
            const unsigned nctas = unsigned((n + 1023) / 1024);
            select_count_kernel<<<nctas, 256, 0, s>>>(r->err, n, B, hist, st, cta_gt, cta_eq);
            if (nctas <= fused_max) {
                select_write_kernel<false><<<nctas, 1024, 0, s>>>(r->err, n, st, cta_gt, cta_eq, sel, hist);
                ctx->launches += 6;
            } else {
                select_scan_kernel<<<1, 1024, 0, s>>>(cta_gt, cta_eq, nctas);
                select_write_kernel<true><<<nctas, 1024, 0, s>>>(r->err, n, st, cta_gt, cta_eq, sel, hist);
                ctx->launches += 7;
            }
        }
        // 2. new sample points of all B splits, one integrand launch
        const uint64_t N = B * Q;
        split_points_kernel<SH, DIM><<<unsigned((N + 255) / 256), 256, 0, s>>>(cap, B, sel, r->rmin, r->rmax, r->errdim, points);
        ctx->launches += 1;
        VB200_CUDA(ctx, cudaGetLastError());
        vb200_eval_launch ev; std::memset(&ev, 0, sizeof(ev));
        ev.n = N; ev.dim = DIM; ev.points = points; ev.values = vals;
        int rc = call_thunk(ctx, f, VB200_K_EVAL_POINTS, &ev); if (rc) return rc;
        // 3. children + heuristics
        if (2 * DIM * Sh::L >= 128 && B <= cta_max)      // few splits of large regions: a CTA per split (the round lasts as long as its slowest split)
            kchild_cta<<<unsigned(B), 256, per_warp, s>>>(cap, B, n, sel, vals, r->rmin, r->rmax, r->data, r->err, r->errdim, heur_args(p->heuristic, p->metric, p->size_weight, p->mixed));
        else
            kchild<<<unsigned((B + wpc - 1) / wpc), wpc * 32, per_warp * wpc, s>>>(cap, B, n, sel, vals, r->rmin, r->rmax, r->data, r->err, r->errdim,
                                                                                   heur_args(p->heuristic, p->metric, p->size_weight, p->mixed));
        ctx->launches++;
        VB200_CUDA(ctx, cudaGetLastError());
        n += B; left -= B;
    }
    r->count = n;
    return VB200_OK;
}

// ---- tolerance-driven refinement (integrator_adaptive_tolerance, reference src/nested/integrator-adaptive-tolerance.h:15-39) --------
// The reference recurses depth first: a region whose heuristic error is below the tolerance is integrated, any other is split and
// its children visited in order.  Whether a region is split depends on that region alone, so the LEAF SET is the same in any
// visiting order: here every round splits all regions that still fail the test at once (same kernels as the batched top-k mode),
// every region carries its root-to-leaf path as a 128-bit key (bit d = which child at depth d), and a final radix sort of the
// left-aligned keys puts the leaves into the reference's depth-first order — the region->bin accumulation that follows then adds
// them in the reference's order, bit for bit.
__global__ void tol_select_kernel(const float* __restrict__ err, const uint32_t* __restrict__ depth, uint64_t n, float tolerance,
                                  unsigned* __restrict__ sel, unsigned* __restrict__ counters /* [0]=count [1]=max depth */) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!(err[i] < tolerance)) {                       // integrator-adaptive-tolerance.h:19 (a NaN error is split, as upstream)
        sel[atomicAdd(&counters[0], 1u)] = unsigned(i);
        atomicMax(&counters[1], depth[i]);
    }
}
__global__ void tol_paths_kernel(const unsigned* __restrict__ sel, uint64_t nsel, uint64_t n_old, unsigned long long* __restrict__ key_hi,
                                 unsigned long long* __restrict__ key_lo, uint32_t* __restrict__ depth) {
    const uint64_t r = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= nsel) return;
    const unsigned slot = sel[r]; const uint64_t slot1 = n_old + r;
    const uint32_t d = depth[slot];
    unsigned long long hi = key_hi[slot], lo = key_lo[slot];
    key_hi[slot1] = d < 64 ? hi | (1ull << (63 - d)) : hi;
    key_lo[slot1] = d < 64 ? lo : lo | (1ull << (127 - d));
    depth[slot] = d + 1; depth[slot1] = d + 1;         // child 0 keeps the parent's bits (a 0 appended), child 1 sets bit d
}
__global__ void tol_iota_kernel(unsigned* idx, uint64_t n) { const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; if (i < n) idx[i] = unsigned(i); }
__global__ void tol_gather_keys_kernel(const unsigned long long* __restrict__ key, const unsigned* __restrict__ idx, unsigned long long* __restrict__ out, uint64_t n) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; if (i < n) out[i] = key[idx[i]];
}
// column gather: dst[k][i] = src[k][perm[i]]   (blockIdx.y strides over the columns)
template<class T>
__global__ void tol_permute_kernel(const T* __restrict__ src, uint64_t src_cap, T* __restrict__ dst, uint64_t dst_cap, const unsigned* __restrict__ perm, uint64_t n, int columns) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t j = perm[i];
    for (int k = blockIdx.y; k < columns; k += gridDim.y) dst[uint64_t(k) * dst_cap + i] = src[uint64_t(k) * src_cap + j];
}

struct TolState {
    vb200_regions* r = nullptr;
    unsigned long long *key_hi = nullptr, *key_lo = nullptr; uint32_t* depth = nullptr;
};

// re-stride the SoA table (and the path columns) to a larger capacity
static int tol_grow(vb200_ctx* ctx, TolState& t, uint64_t n, uint64_t new_cap) {
    vb200_regions* o = t.r; vb200_regions* g = nullptr;
    int rc = regions_alloc(ctx, o->dim, o->rule, new_cap, &g); if (rc) return rc;
    unsigned long long *kh = nullptr, *kl = nullptr; uint32_t* dp = nullptr;
    if (dmalloc(ctx, &kh, new_cap * 8) != cudaSuccess || dmalloc(ctx, &kl, new_cap * 8) != cudaSuccess || dmalloc(ctx, &dp, new_cap * 4) != cudaSuccess) {
        cudaGetLastError(); dfree(ctx, kh); dfree(ctx, kl); dfree(ctx, dp); vb200_regions_free(g);
        return fail(ctx, VB200_ERR_NOMEM, "tolerance refinement: %llu region slots do not fit", (unsigned long long)new_cap);
    }
    cudaStream_t s = ctx->stream;
    const size_t w = n * sizeof(float);
    VB200_CUDA(ctx, cudaMemcpy2DAsync(g->rmin, new_cap * 4, o->rmin, o->capacity * 4, w, o->dim, cudaMemcpyDeviceToDevice, s));
    VB200_CUDA(ctx, cudaMemcpy2DAsync(g->rmax, new_cap * 4, o->rmax, o->capacity * 4, w, o->dim, cudaMemcpyDeviceToDevice, s));
    VB200_CUDA(ctx, cudaMemcpy2DAsync(g->data, new_cap * 4, o->data, o->capacity * 4, w, o->sd, cudaMemcpyDeviceToDevice, s));
    VB200_CUDA(ctx, cudaMemcpyAsync(g->err, o->err, w, cudaMemcpyDeviceToDevice, s));
    VB200_CUDA(ctx, cudaMemcpyAsync(g->errdim, o->errdim, n * 4, cudaMemcpyDeviceToDevice, s));
    VB200_CUDA(ctx, cudaMemcpyAsync(kh, t.key_hi, n * 8, cudaMemcpyDeviceToDevice, s));
    VB200_CUDA(ctx, cudaMemcpyAsync(kl, t.key_lo, n * 8, cudaMemcpyDeviceToDevice, s));
    VB200_CUDA(ctx, cudaMemcpyAsync(dp, t.depth, n * 4, cudaMemcpyDeviceToDevice, s));
    dfree(ctx, t.key_hi); dfree(ctx, t.key_lo); dfree(ctx, t.depth); vb200_regions_free(o);
    t.r = g; t.key_hi = kh; t.key_lo = kl; t.depth = dp;
    return VB200_OK;
}

template<int SH, int SL, int DIM>
int run_tolerance(vb200_ctx* ctx, const vb200_integrand* f, const vb200_tolerance_params* p, TolState& t, uint64_t* n_out) {
    using Sh = D_::GreedyShape<SH, SL, DIM>;
    cudaStream_t s = ctx->stream;
    root_error_kernel<SH, SL, DIM><<<1, 32, (Sh::SD + Sh::L + 4 * DIM) * sizeof(float), s>>>(t.r->capacity, t.r->rmin, t.r->rmax, t.r->data, t.r->err, t.r->errdim,
            heur_args(p->heuristic, p->metric, p->size_weight, vb200_mixed_heuristic{}));
    ctx->launches++;
    VB200_CUDA(ctx, cudaGetLastError());
    const size_t per_warp = (3 * Sh::SD + 2 * DIM * Sh::L + 4 * DIM + 2 * DIM + 2) * sizeof(float);
    int wpc = int((96u << 10) / per_warp); if (wpc > 8) wpc = 8; if (wpc < 1) wpc = 1;
    auto kchild = split_children_kernel<SH, SL, DIM, false>;
    auto kchild_cta = split_children_kernel<SH, SL, DIM, true>;
    VB200_CUDA(ctx, cudaFuncSetAttribute(kchild, cudaFuncAttributeMaxDynamicSharedMemorySize, int(per_warp * wpc)));
    VB200_CUDA(ctx, cudaFuncSetAttribute(kchild_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, int(per_warp)));
    const uint64_t cta_max = split_cta_max();
    const uint64_t Q = uint64_t(SH - 1) * Sh::L;
    uint64_t max_batch = (32ull << 20) / Q; if (max_batch < 1) max_batch = 1;
    const uint64_t limit = p->max_regions ? p->max_regions : (1ull << 27);
    unsigned* counters = nullptr;
    if (dmalloc(ctx, &counters, 2 * sizeof(unsigned)) != cudaSuccess) { cudaGetLastError(); return fail(ctx, VB200_ERR_NOMEM, "out of device memory"); }
    uint64_t n = 1; int rc = VB200_OK;
    for (;;) {
        unsigned* sel = nullptr;
        if (dmalloc(ctx, &sel, n * sizeof(unsigned)) != cudaSuccess) { cudaGetLastError(); rc = fail(ctx, VB200_ERR_NOMEM, "out of device memory"); break; }
        unsigned h[2] = {0, 0};
        if (cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned), s) != cudaSuccess) { dfree(ctx, sel); rc = fail(ctx, VB200_ERR_CUDA, "memset failed"); break; }
        tol_select_kernel<<<unsigned((n + 255) / 256), 256, 0, s>>>(t.r->err, t.depth, n, p->tolerance, sel, counters);
        ctx->launches++;
        if (cudaMemcpyAsync(h, counters, sizeof(h), cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
            dfree(ctx, sel); rc = fail(ctx, VB200_ERR_CUDA, "tolerance refinement failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
        const uint64_t nsel = h[0];
        if (nsel == 0) { dfree(ctx, sel); break; }
        if (h[1] >= 128) { dfree(ctx, sel); rc = fail(ctx, VB200_ERR_UNSUPPORTED,
                "tolerance %g not reached after 128 levels of subdivision (the reference would recurse without bound)", double(p->tolerance)); break; }
        if (n + nsel > limit) { dfree(ctx, sel); rc = fail(ctx, VB200_ERR_NOMEM, "tolerance refinement exceeds %llu regions", (unsigned long long)limit); break; }
        if (n + nsel > t.r->capacity) { rc = tol_grow(ctx, t, n, std::max<uint64_t>(2 * t.r->capacity, n + nsel)); if (rc) { dfree(ctx, sel); break; } }
        const uint64_t cap = t.r->capacity;
        for (uint64_t off = 0; off < nsel && !rc; off += max_batch) {
            const uint64_t B = std::min<uint64_t>(max_batch, nsel - off), N = B * Q;
            float *points = nullptr, *vals = nullptr;
            if (dmalloc(ctx, &points, N * DIM * sizeof(float)) != cudaSuccess || dmalloc(ctx, &vals, N * sizeof(float)) != cudaSuccess) {
                cudaGetLastError(); dfree(ctx, points); dfree(ctx, vals); rc = fail(ctx, VB200_ERR_NOMEM, "out of device memory"); break; }
            split_points_kernel<SH, DIM><<<unsigned((N + 255) / 256), 256, 0, s>>>(cap, B, sel + off, t.r->rmin, t.r->rmax, t.r->errdim, points);
            tol_paths_kernel<<<unsigned((B + 255) / 256), 256, 0, s>>>(sel + off, B, n + off, t.key_hi, t.key_lo, t.depth);
            ctx->launches += 2;
            vb200_eval_launch ev; std::memset(&ev, 0, sizeof(ev));
            ev.n = N; ev.dim = DIM; ev.points = points; ev.values = vals;
            rc = call_thunk(ctx, f, VB200_K_EVAL_POINTS, &ev);
            if (!rc) {
                if (2 * DIM * Sh::L >= 128 && B <= cta_max)
                    kchild_cta<<<unsigned(B), 256, per_warp, s>>>(cap, B, n + off, sel + off, vals, t.r->rmin, t.r->rmax, t.r->data, t.r->err, t.r->errdim,
                                                                  heur_args(p->heuristic, p->metric, p->size_weight, vb200_mixed_heuristic{}));
                else
                    kchild<<<unsigned((B + wpc - 1) / wpc), wpc * 32, per_warp * wpc, s>>>(cap, B, n + off, sel + off, vals, t.r->rmin, t.r->rmax, t.r->data, t.r->err, t.r->errdim,
                                                                                           heur_args(p->heuristic, p->metric, p->size_weight, vb200_mixed_heuristic{}));
                ctx->launches++;
                if (cudaGetLastError() != cudaSuccess) rc = fail(ctx, VB200_ERR_CUDA, "split kernel launch failed");
            }
            dfree(ctx, points); dfree(ctx, vals);
        }
        dfree(ctx, sel);
        if (rc) break;
        n += nsel;
    }
    dfree(ctx, counters);
    *n_out = n;
    return rc;
}

} // namespace

namespace vb200 {

__global__ void root_points_kernel(int S, int dim, uint64_t n, const float* rmin, const float* rmax, float* points) {
    const uint64_t k = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint64_t t = k;
    for (int d = 0; d < dim; ++d) {
        const double p = R::dd(double(t % uint64_t(S)), double(S - 1)); t /= uint64_t(S);
        points[uint64_t(d) * n + k] = D_::grid_coord(p, rmin[d], rmax[d]);
    }
}
__global__ void scatter_root_kernel(uint64_t n, uint64_t cap, int dim, const float* vals, const float* lo, const float* hi, float* data, float* rmin, float* rmax) {
    const uint64_t k = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k < n) data[k * cap] = vals[k];
    if (k < uint64_t(dim)) { rmin[k * cap] = lo[k]; rmax[k * cap] = hi[k]; }
}

int generate_batched(vb200_ctx* ctx, const vb200_integrand* f, const vb200_adaptive_params* p, vb200_regions** out) {
    vb200_regions* r = nullptr;
    const uint64_t cap = p->iterations + 1;
    int rc = regions_alloc(ctx, f->dim, p->rule, cap, &r); if (rc) return rc;
    const int D = f->dim, S = r->SH; const uint64_t sd = uint64_t(r->sd);
    uint64_t L = 1; for (int i = 0; i < D - 1; ++i) L *= uint64_t(S);
    const uint64_t Q = uint64_t(S - 1) * L;
    // bound the per-round buffers: at most ~32 Mi new sample points per round
    uint64_t max_batch = (32ull << 20) / Q; if (max_batch < 1) max_batch = 1;
    const uint64_t maxN = std::max<uint64_t>(std::min<uint64_t>(max_batch, cap) * Q, sd);
    SelectState* st = nullptr; unsigned *hist = nullptr, *cta_gt = nullptr, *cta_eq = nullptr, *sel = nullptr; float *points = nullptr, *vals = nullptr, *lohi = nullptr;
    auto cleanup = [&] () { dfree(ctx, st); dfree(ctx, hist); dfree(ctx, cta_gt); dfree(ctx, cta_eq); dfree(ctx, sel); dfree(ctx, points); dfree(ctx, vals); dfree(ctx, lohi); };
    auto bail = [&] (int code) { cudaStreamSynchronize(ctx->stream); cleanup(); vb200_regions_free(r); return code; };
    const uint64_t nctas_max = (cap + 1023) / 1024;
    if (dmalloc(ctx, &st, sizeof(SelectState)) != cudaSuccess || dmalloc(ctx, &hist, 4 * 256 * sizeof(unsigned)) != cudaSuccess ||
        dmalloc(ctx, &cta_gt, nctas_max * sizeof(unsigned)) != cudaSuccess || dmalloc(ctx, &cta_eq, nctas_max * sizeof(unsigned)) != cudaSuccess ||
        dmalloc(ctx, &sel, std::min<uint64_t>(max_batch, cap) * sizeof(unsigned)) != cudaSuccess || dmalloc(ctx, &points, maxN * D * sizeof(float)) != cudaSuccess ||
        dmalloc(ctx, &vals, maxN * sizeof(float)) != cudaSuccess || dmalloc(ctx, &lohi, 2 * VB200_MAX_DIM * sizeof(float)) != cudaSuccess) {
        cudaGetLastError(); return bail(fail(ctx, VB200_ERR_NOMEM, "working set of the batched refinement does not fit")); }
    // root region (regions-generator-adaptive-heap.h:27)
    float h_lohi[2 * VB200_MAX_DIM];
    for (int d = 0; d < D; ++d) { h_lohi[d] = p->domain.rmin[d]; h_lohi[VB200_MAX_DIM + d] = p->domain.rmax[d]; }
    if (cudaMemcpyAsync(lohi, h_lohi, sizeof(h_lohi), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return bail(fail(ctx, VB200_ERR_CUDA, "range upload failed"));
    root_points_kernel<<<unsigned((sd + 255) / 256), 256, 0, ctx->stream>>>(S, D, sd, lohi, lohi + VB200_MAX_DIM, points);
    ctx->launches++;
    vb200_eval_launch ev; std::memset(&ev, 0, sizeof(ev));
    ev.n = sd; ev.dim = D; ev.points = points; ev.values = vals;
    rc = call_thunk(ctx, f, VB200_K_EVAL_POINTS, &ev); if (rc) return bail(rc);
    scatter_root_kernel<<<unsigned((std::max<uint64_t>(sd, uint64_t(D)) + 255) / 256), 256, 0, ctx->stream>>>(sd, cap, D, vals, lohi, lohi + VB200_MAX_DIM, r->data, r->rmin, r->rmax);
    ctx->launches++;
    rc = VB200_ERR_UNSUPPORTED;
#define VB200_RB(SH_, SL_, DD) if (r->SH == SH_ && r->SL == SL_ && D == DD) rc = run_rounds<SH_, SL_, DD>(ctx, f, p, r, st, hist, cta_gt, cta_eq, sel, points, vals, max_batch);
    VB200_RB(3, 2, 1) VB200_RB(3, 2, 2) VB200_RB(3, 2, 3) VB200_RB(3, 2, 4) VB200_RB(3, 2, 5) VB200_RB(3, 2, 6)
    VB200_RB(5, 3, 1) VB200_RB(5, 3, 2) VB200_RB(5, 3, 3) VB200_RB(5, 3, 4) VB200_RB(5, 3, 5)
#undef VB200_RB
    if (rc == VB200_ERR_UNSUPPORTED) return bail(fail(ctx, VB200_ERR_UNSUPPORTED, "batched refinement: rule/dimension combination not instantiated"));
    if (rc) return bail(rc);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return bail(fail(ctx, VB200_ERR_CUDA, "batched refinement failed: %s", cudaGetErrorString(e)));
    cleanup();
    *out = r;
    return VB200_OK;
}

int generate_tolerance(vb200_ctx* ctx, const vb200_integrand* f, const vb200_tolerance_params* p, vb200_regions** out) {
    const int D = f->dim;
    TolState t;
    uint64_t cap = 4096;
    int rc = regions_alloc(ctx, D, p->rule, cap, &t.r); if (rc) return rc;
    const int S = t.r->SH; const uint64_t sd = uint64_t(t.r->sd);
    float *points = nullptr, *vals = nullptr, *lohi = nullptr;
    auto cleanup = [&] () { dfree(ctx, points); dfree(ctx, vals); dfree(ctx, lohi); dfree(ctx, t.key_hi); dfree(ctx, t.key_lo); dfree(ctx, t.depth);
            t.key_hi = t.key_lo = nullptr; t.depth = nullptr; };
    auto bail = [&] (int code) { cudaStreamSynchronize(ctx->stream); cleanup(); vb200_regions_free(t.r); return code; };
    if (dmalloc(ctx, &t.key_hi, cap * 8) != cudaSuccess || dmalloc(ctx, &t.key_lo, cap * 8) != cudaSuccess || dmalloc(ctx, &t.depth, cap * 4) != cudaSuccess ||
        dmalloc(ctx, &points, sd * D * sizeof(float)) != cudaSuccess || dmalloc(ctx, &vals, sd * sizeof(float)) != cudaSuccess || dmalloc(ctx, &lohi,
                2 * VB200_MAX_DIM * sizeof(float)) != cudaSuccess) {
        cudaGetLastError(); return bail(fail(ctx, VB200_ERR_NOMEM, "working set of the tolerance refinement does not fit")); }
    if (cudaMemsetAsync(t.key_hi, 0, 8, ctx->stream) != cudaSuccess || cudaMemsetAsync(t.key_lo, 0, 8, ctx->stream) != cudaSuccess || cudaMemsetAsync(t.depth, 0, 4, ctx->stream) != cudaSuccess)
        return bail(fail(ctx, VB200_ERR_CUDA, "memset failed"));
    // root region (integrator-adaptive-tolerance.h:37)
    float h_lohi[2 * VB200_MAX_DIM];
    for (int d = 0; d < D; ++d) { h_lohi[d] = p->domain.rmin[d]; h_lohi[VB200_MAX_DIM + d] = p->domain.rmax[d]; }
    if (cudaMemcpyAsync(lohi, h_lohi, sizeof(h_lohi), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return bail(fail(ctx, VB200_ERR_CUDA, "range upload failed"));
    root_points_kernel<<<unsigned((sd + 255) / 256), 256, 0, ctx->stream>>>(S, D, sd, lohi, lohi + VB200_MAX_DIM, points);
    ctx->launches++;
    vb200_eval_launch ev; std::memset(&ev, 0, sizeof(ev));
    ev.n = sd; ev.dim = D; ev.points = points; ev.values = vals;
    rc = call_thunk(ctx, f, VB200_K_EVAL_POINTS, &ev); if (rc) return bail(rc);
    scatter_root_kernel<<<unsigned((std::max<uint64_t>(sd, uint64_t(D)) + 255) / 256), 256, 0, ctx->stream>>>(sd, cap, D, vals, lohi, lohi + VB200_MAX_DIM, t.r->data, t.r->rmin, t.r->rmax);
    ctx->launches++;
    uint64_t n = 0;
    rc = VB200_ERR_UNSUPPORTED;
#define VB200_RT(SH_, SL_, DD) if (t.r->SH == SH_ && t.r->SL == SL_ && D == DD) rc = run_tolerance<SH_, SL_, DD>(ctx, f, p, t, &n);
    VB200_RT(3, 2, 1) VB200_RT(3, 2, 2) VB200_RT(3, 2, 3) VB200_RT(3, 2, 4) VB200_RT(3, 2, 5) VB200_RT(3, 2, 6)
    VB200_RT(5, 3, 1) VB200_RT(5, 3, 2) VB200_RT(5, 3, 3) VB200_RT(5, 3, 4) VB200_RT(5, 3, 5)
#undef VB200_RT
    if (rc == VB200_ERR_UNSUPPORTED && n == 0) return bail(fail(ctx, VB200_ERR_UNSUPPORTED, "tolerance refinement: rule/dimension combination not instantiated"));
    if (rc) return bail(rc);
    // leaves -> the reference's depth-first order: stable LSD radix sort of the 128-bit path keys (low word, then high word)
    cudaStream_t s = ctx->stream;
    unsigned *idx_a = nullptr, *idx_b = nullptr; unsigned long long *k_in = nullptr, *k_out = nullptr; void* tmp = nullptr; size_t tmp_bytes = 0;
    vb200_regions* fin = nullptr;
    auto cleanup2 = [&] () { dfree(ctx, idx_a); dfree(ctx, idx_b); dfree(ctx, k_in); dfree(ctx, k_out); dfree(ctx, tmp); };
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, idx_a, idx_b, int(n), 0, 64, s);
    if (dmalloc(ctx, &idx_a, n * 4) != cudaSuccess || dmalloc(ctx, &idx_b, n * 4) != cudaSuccess || dmalloc(ctx, &k_in, n * 8) != cudaSuccess ||
        dmalloc(ctx, &k_out, n * 8) != cudaSuccess || dmalloc_bytes(ctx, &tmp, tmp_bytes) != cudaSuccess) { cudaGetLastError(); cleanup2(); return bail(fail(ctx,
                VB200_ERR_NOMEM, "out of device memory")); }
    const unsigned g1 = unsigned((n + 255) / 256);
    tol_iota_kernel<<<g1, 256, 0, s>>>(idx_a, n);
    cudaMemcpyAsync(k_in, t.key_lo, n * 8, cudaMemcpyDeviceToDevice, s);
    cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, idx_a, idx_b, int(n), 0, 64, s);
    tol_gather_keys_kernel<<<g1, 256, 0, s>>>(t.key_hi, idx_b, k_in, n);
    cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, idx_b, idx_a, int(n), 0, 64, s);
    ctx->launches += 4;
    rc = regions_alloc(ctx, D, p->rule, n, &fin); if (rc) { cleanup2(); return bail(rc); }
    const uint64_t ocap = t.r->capacity;
    tol_permute_kernel<float><<<dim3(g1, unsigned(std::min<uint64_t>(sd, 64))), 256, 0, s>>>(t.r->data, ocap, fin->data, n, idx_a, n, int(sd));
    tol_permute_kernel<float><<<dim3(g1, unsigned(D)), 256, 0, s>>>(t.r->rmin, ocap, fin->rmin, n, idx_a, n, D);
    tol_permute_kernel<float><<<dim3(g1, unsigned(D)), 256, 0, s>>>(t.r->rmax, ocap, fin->rmax, n, idx_a, n, D);
    tol_permute_kernel<float><<<dim3(g1, 1), 256, 0, s>>>(t.r->err, ocap, fin->err, n, idx_a, n, 1);
    tol_permute_kernel<uint32_t><<<dim3(g1, 1), 256, 0, s>>>(t.r->errdim, ocap, fin->errdim, n, idx_a, n, 1);
    ctx->launches += 5;
    cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cleanup2();
    if (e != cudaSuccess) { vb200_regions_free(fin); return bail(fail(ctx, VB200_ERR_CUDA, "tolerance refinement (ordering pass) failed: %s", cudaGetErrorString(e))); }
    cleanup(); vb200_regions_free(t.r);
    fin->count = n;
    *out = fin;
    return VB200_OK;
}

} // namespace vb200
