// K3 — infinite-dimensional (RangeInfinite) per-bin Monte Carlo, replacing the reference's
//   MonteCarloPerBinParallel::integrate(RangeInfinite)   src/monte-carlo/monte-carlo-per-bin-parallel.h:73-100
//   RandomSequenceRefDis (lazy sequence over the bin RNG)  src/monte-carlo/random-sequence-ref-dis.h:11-44
//   RangeInfinite (implicit [0,1] tail)                    src/range-infinite.h:16-64
// The integrand is a functor over a *sequence* (seq.begin(), *it, ++it; never-ending), exactly the reference's
// protocol.  On the GPU the sequence is stateless: element i of sample s in bin b is
//   u01(Philox4x32-10(key=seed, counter=(b, s, i/4))[i%4]) * (max_i - min_i) + min_i
// so a lane needs no per-bin generator state and any lane can produce any element.  Each lane owns whole paths
// and runs its own Russian roulette inside the user's functor.
#pragma once
#include <type_traits>
#include <cuda_runtime.h>
#include <math_constants.h>
#include "../../viltrum_b200.h"
#include "philox.cuh"
#include "mc_per_bin.cuh"

namespace viltrum { namespace b200 { namespace device {

template<int DIMBINS>
struct PhiloxSequence {
    uint32_t b0, b1, s, k0, k1;
    float lo[DIMBINS], ext[DIMBINS];        // the bin's box in the binned dimensions
    const vb200_domain* dom;                 // explicit entries beyond the binned dims (kernel parameter space)

    class const_iterator {
        const PhiloxSequence* q; uint32_t i; u32x4 blk; float n;
        __device__ __forceinline__ void load() {
            if ((i & 3u) == 0u) blk = philox4x32<10>(u32x4{q->b0, q->b1, q->s, i >> 2}, q->k0, q->k1);
            const float u = pick(blk, int(i & 3u));
            float v = u;                                                   // default range [0,1): u*(1-0)+0 == u
            bool binned = false;
#pragma unroll
            for (int d = 0; d < DIMBINS; ++d) if (i == uint32_t(d)) { v = fmaf(u, q->ext[d], q->lo[d]); binned = true; }
            if (!binned && int(i) < q->dom->dim) v = fmaf(u, q->dom->rmax[i] - q->dom->rmin[i], q->dom->rmin[i]);
            n = v;
        }
    public:
        __device__ __forceinline__ const_iterator(const PhiloxSequence* q_) : q(q_), i(0) { load(); }
        __device__ __forceinline__ const_iterator(const PhiloxSequence* q_, int) : q(q_), i(0), blk{0, 0, 0, 0}, n(0.0f) {}      // unarmed (wavefront kernel)
        __device__ __forceinline__ const float& operator*() const { return n; }
        __device__ __forceinline__ const_iterator& operator++() { ++i; load(); return *this; }
        __device__ __forceinline__ bool operator!=(const const_iterator&) const { return true; }   // infinite list
        __device__ __forceinline__ bool operator==(const const_iterator&) const { return false; }
    };
    __device__ __forceinline__ const_iterator begin() const { return const_iterator(this); }
    __device__ __forceinline__ const_iterator end() const { return const_iterator(this); }
};

// returns RangeInfinite::_volume of the bin sub-range: float product over its explicit entries, in order (range-infinite.h:22-23)
template<int DIMBINS>
__device__ __forceinline__ float walk_bin_box(const vb200_domain& dom, uint64_t bin, float (&lo)[DIMBINS], float (&ext)[DIMBINS]) {
    uint32_t pos[VB200_MAX_DIMBINS];
    unflatten_bin<DIMBINS>(bin, dom, pos);
#pragma unroll
    for (int i = 0; i < DIMBINS; ++i) {
        // RangeInfinite::min/max default to 0/1 beyond the explicit entries (range-infinite.h:31-37)
        const float rmin = i < dom.dim ? dom.rmin[i] : 0.0f;
        const float drange = dom.drange[i];
        const float a = __fadd_rn(rmin, __fmul_rn(float(pos[i]), drange));
        const float b = __fadd_rn(rmin, __fmul_rn(float(pos[i] + 1u), drange));
        lo[i] = a; ext[i] = b - a;
    }
    float volume = 1.0f;
#pragma unroll
    for (int i = 0; i < DIMBINS; ++i) volume = __fmul_rn(volume, ext[i]);
    for (int i = DIMBINS; i < dom.dim; ++i) volume = __fmul_rn(volume, dom.rmax[i] - dom.rmin[i]);
    return volume;
}

// value written for a bin: flavor 0  sum f * vol(range)/spp                       (monte-carlo-per-bin-parallel.h:77,96)
//                          flavor 1  nbins * float(sum f * (vol(bin box)/spp))      (monte-carlo.h:70-72,81; integrator-per-bin-parallel.h:33)
__device__ __forceinline__ float walk_bin_value(const vb200_walk_launch& a, float sum, float volume) {
    if (a.flavor == VB200_PER_BIN_MC) {
        const float sol = float(double(sum) * (double(volume) / double(a.spp)));
        return float(double(a.nbins_total) * double(sol));
    }
    return float(double(sum) * a.factor);
}

template<class F, int DIMBINS, bool MOMENTS, bool EXACT>
__global__ void __launch_bounds__(MC_THREADS)
walk_kernel(const F f, const vb200_walk_launch a) {
    const uint32_t LPB = a.lanes_per_bin, G = 32u / LPB;
    const uint32_t lane = threadIdx.x & 31u, sub = lane % LPB, grp = lane / LPB;
    const uint64_t nshard = a.bin_end - a.bin_begin;
    const uint64_t ntiles = (nshard + G - 1) / G;
    uint64_t tile = 0;
    if (lane == 0) tile = atomicAdd(a.tile_counter, 1ull);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    while (tile < ntiles) {
        const uint64_t bin = a.bin_begin + tile * G + grp;
        const bool live = bin < a.bin_end;
        float sum = 0.0f, sum2 = 0.0f, volume = 1.0f;
        if (live) {
            PhiloxSequence<DIMBINS> seq;
            seq.b0 = uint32_t(bin); seq.b1 = uint32_t(bin >> 32); seq.k0 = a.key0; seq.k1 = a.key1; seq.dom = &a.domain;
            volume = walk_bin_box<DIMBINS>(a.domain, bin, seq.lo, seq.ext);
            for (uint32_t s = sub; s < a.spp; s += LPB) {
                seq.s = s;
                const float v = f(seq);
                sum += v;
                if (MOMENTS) sum2 = fmaf(v, v, sum2);
            }
        }
        for (uint32_t off = LPB >> 1; off > 0; off >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, off);
            if (MOMENTS) sum2 += __shfl_xor_sync(0xffffffffu, sum2, off);
        }
        if (live && sub == 0) {
            const float v = walk_bin_value(a, sum, volume);
            a.out[bin] = a.accumulate ? float(double(a.out[bin]) + double(v)) : v;
            if (MOMENTS) {
                if (a.sum_f)  a.sum_f[bin - a.bin_begin]  = sum;
                if (a.sum_f2) a.sum_f2[bin - a.bin_begin] = sum2;
            }
        }
        signal_tile_done(a.signal, tile, ntiles, lane);
        if (lane == 0) tile = atomicAdd(a.tile_counter, 1ull);
        tile = __shfl_sync(0xffffffffu, tile, 0);
    }
}

// ---- wavefront variant: per-lane Russian roulette with lane refill ---------------------------------------------------------
// A functor may ALSO describe itself as a state machine (besides the reference's operator()(seq)):
//     struct State {...};
//     template<class It> __device__ State begin(It& it) const;          // start of a path: consumes its first elements
//     template<class It> __device__ bool  step(State& s, It& it) const; // one roulette round; false = the path ended
//     __device__ float end(const State& s) const;                       // the path's value
// with begin/step/end performing exactly the arithmetic of operator().  Then a lane whose path has ended does not idle until the
// longest path of its warp is done: dead lanes are re-armed with their next sample as soon as REFILL of the 32 lanes are dead,
// and all live lanes execute step() together.  Paths consume the same Philox elements in the same order as in the generic kernel
// and a lane adds its samples in the same order, so the bins are bit-identical to walk_kernel's.
template<class F, class = void> struct has_steps : std::false_type {};
template<class F> struct has_steps<F, std::void_t<typename F::State>> : std::true_type {};

// REFILL: dead lanes needed before a re-arm round; STEPS: roulette rounds per loop iteration.  A path consumes a fixed number of
// sequence elements per begin()/step(); when STEPS*elements-per-step is a multiple of 4 every live lane crosses its Philox block
// boundary at the same instruction, so the generator runs once per iteration at full lane utilisation instead of once per step
// at half (measured: profiles/walk_variants_r1.txt).
template<class F, int DIMBINS, bool MOMENTS, bool EXACT, int REFILL = 8, int STEPS = 2>
__global__ void __launch_bounds__(MC_THREADS)
walk_wavefront_kernel(const F f, const vb200_walk_launch a) {
    const uint32_t LPB = a.lanes_per_bin, G = 32u / LPB;
    const uint32_t lane = threadIdx.x & 31u, sub = lane % LPB, grp = lane / LPB;
    const uint64_t nshard = a.bin_end - a.bin_begin;
    const uint64_t ntiles = (nshard + G - 1) / G;
    uint64_t tile = 0;
    if (lane == 0) tile = atomicAdd(a.tile_counter, 1ull);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    while (tile < ntiles) {
        const uint64_t bin = a.bin_begin + tile * G + grp;
        const bool live = bin < a.bin_end;
        float sum = 0.0f, sum2 = 0.0f, volume = 1.0f;
        PhiloxSequence<DIMBINS> seq;
        seq.b0 = uint32_t(bin); seq.b1 = uint32_t(bin >> 32); seq.k0 = a.key0; seq.k1 = a.key1; seq.dom = &a.domain; seq.s = 0;
        if (live) volume = walk_bin_box<DIMBINS>(a.domain, bin, seq.lo, seq.ext);
        else { for (int d = 0; d < DIMBINS; ++d) { seq.lo[d] = 0.0f; seq.ext[d] = 1.0f; } }
        uint32_t next = live ? sub : a.spp;          // next sample this lane will start
        bool alive = false;
        typename F::State st;
        typename PhiloxSequence<DIMBINS>::const_iterator it(&seq, 0);      // unarmed
        while (true) {
            const unsigned alive_mask = __ballot_sync(0xffffffffu, alive);
            const unsigned want_mask = __ballot_sync(0xffffffffu, !alive && next < a.spp);
            if (alive_mask == 0u && want_mask == 0u) break;
            // re-arm dead lanes in batches: when enough of them wait, or when nobody is alive any more
            if (want_mask != 0u && (__popc(want_mask) >= REFILL || alive_mask == 0u)) {
                if (!alive && next < a.spp) {
                    seq.s = next; next += LPB;
                    it = seq.begin();
                    st = f.begin(it);
                    alive = true;
                }
            }
#pragma unroll
            for (int rep = 0; rep < STEPS; ++rep) {
                if (alive) {
                    if (!f.step(st, it)) {
                        const float v = f.end(st);
                        sum += v;
                        if (MOMENTS) sum2 = fmaf(v, v, sum2);
                        alive = false;
                    }
                }
            }
        }
        for (uint32_t off = LPB >> 1; off > 0; off >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, off);
            if (MOMENTS) sum2 += __shfl_xor_sync(0xffffffffu, sum2, off);
        }
        if (live && sub == 0) {
            const float v = walk_bin_value(a, sum, volume);
            a.out[bin] = a.accumulate ? float(double(a.out[bin]) + double(v)) : v;
            if (MOMENTS) {
                if (a.sum_f)  a.sum_f[bin - a.bin_begin]  = sum;
                if (a.sum_f2) a.sum_f2[bin - a.bin_begin] = sum2;
            }
        }
        signal_tile_done(a.signal, tile, ntiles, lane);
        if (lane == 0) tile = atomicAdd(a.tile_counter, 1ull);
        tile = __shfl_sync(0xffffffffu, tile, 0);
    }
}

// ---- block-fed wavefront variant: ONE Philox call per lane and loop iteration, immediate refill -----------------------------------
// A state-machine functor may also declare how many sequence elements its parts consume:
//     static constexpr int elements_begin = 2;    // begin() reads exactly this many (2, or 0)
//     static constexpr int elements_step  = 2;    // step() reads at most this many (fewer only when it returns false)
// With 2 + 2 the element stream of a path falls into Philox blocks the kernel can hand out whole: block 0 = begin() + one roulette round,
// every later block = two rounds (with 0 + 2 every block is two rounds).  Every lane then draws exactly one block per loop iteration — the generator (the expensive part:
// 20 IMAD.WIDE) runs convergent at full lane utilisation — a lane whose path ended is re-armed with its next sample in the very next
// iteration (no batching needed: begin() costs no extra generator call), and the functor reads its elements through an iterator over
// four registers instead of the general PhiloxSequence iterator (no per-element block/range bookkeeping).  Elements are the same
// Philox words mapped with the same expressions, paths are summed per lane in the same order: bins are bit-identical to walk_kernel's.
// Requires every explicit range entry to sit in block 0 (domain.dim <= 4, checked by the launcher).
template<class F, class = void> struct has_block_steps : std::false_type {};
template<class F> struct has_block_steps<F, std::void_t<typename F::State, decltype(F::elements_begin), decltype(F::elements_step)>>
    : std::integral_constant<bool, (F::elements_begin == 2 || F::elements_begin == 0) && F::elements_step == 2> {};

struct BlockIterator {          // the iterator protocol of the reference's sequences (*it, ++it) over one Philox block held in registers
    float e0, e1, e2, e3; int i;
    __device__ __forceinline__ float operator*() const { return i == 0 ? e0 : i == 1 ? e1 : i == 2 ? e2 : e3; }
    __device__ __forceinline__ BlockIterator& operator++() { ++i; return *this; }
};

template<class F, int DIMBINS, bool MOMENTS, bool EXACT>
__global__ void __launch_bounds__(MC_THREADS)
walk_block_kernel(const F f, const vb200_walk_launch a) {
    const uint32_t LPB = a.lanes_per_bin, G = 32u / LPB;
    const uint32_t lane = threadIdx.x & 31u, sub = lane % LPB, grp = lane / LPB;
    const uint64_t nshard = a.bin_end - a.bin_begin;
    const uint64_t ntiles = (nshard + G - 1) / G;
    uint64_t tile = 0;
    if (lane == 0) tile = atomicAdd(a.tile_counter, 1ull);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    while (tile < ntiles) {
        const uint64_t bin = a.bin_begin + tile * G + grp;
        const bool live = bin < a.bin_end;
        float sum = 0.0f, sum2 = 0.0f, volume = 1.0f;
        // element -> value map of block 0: v = fmaf(u, scale, offset); binned dimensions use the bin box, other explicit entries their
        // range, everything else [0,1) (scale 1, offset 0: fmaf(u,1,0) == u)
        float sc[4] = {1.0f, 1.0f, 1.0f, 1.0f}, of[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (live) {
            float lo[DIMBINS], ext[DIMBINS];
            volume = walk_bin_box<DIMBINS>(a.domain, bin, lo, ext);
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                if (d < DIMBINS) { sc[d] = ext[d]; of[d] = lo[d]; }
                else if (d < a.domain.dim) { sc[d] = a.domain.rmax[d] - a.domain.rmin[d]; of[d] = a.domain.rmin[d]; }
            }
        }
        const uint32_t b0 = uint32_t(bin), b1 = uint32_t(bin >> 32);
        uint32_t next = live ? sub : a.spp;          // next sample this lane will start
        uint32_t s = 0, blk = 0;
        bool alive = false;
        typename F::State st;
        while (true) {
            const bool starting = !alive && next < a.spp;
            if (!__any_sync(0xffffffffu, alive || starting)) break;
            if (starting) { s = next; next += LPB; blk = 0; }
            const u32x4 r = philox4x32<10>(u32x4{b0, b1, s, blk}, a.key0, a.key1);
            BlockIterator it;
            it.e0 = u01(r.x); it.e1 = u01(r.y); it.e2 = u01(r.z); it.e3 = u01(r.w); it.i = 0;
            if (blk == 0) {       // PhiloxSequence::const_iterator::load: fmaf(u, extent, lower)
                it.e0 = fmaf(it.e0, sc[0], of[0]); it.e1 = fmaf(it.e1, sc[1], of[1]); it.e2 = fmaf(it.e2, sc[2], of[2]); it.e3 = fmaf(it.e3, sc[3], of[3]);
            }
            ++blk;
            bool ended = false;
            if (starting) {                                              // elements 0,1 — or none, then the first round takes them
                st = f.begin(it); alive = true;
                if constexpr (F::elements_begin == 0) { if (!f.step(st, it)) ended = true; }
            }
            else if (alive) { if (!f.step(st, it)) ended = true; }       // elements 0,1 (or only 0)
            it.i = 2;
            if (alive && !ended) { if (!f.step(st, it)) ended = true; }  // elements 2,3 (or only 2)
            if (ended) {
                const float v = f.end(st);
                sum += v;
                if (MOMENTS) sum2 = fmaf(v, v, sum2);
                alive = false;
            }
        }
        for (uint32_t off = LPB >> 1; off > 0; off >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, off);
            if (MOMENTS) sum2 += __shfl_xor_sync(0xffffffffu, sum2, off);
        }
        if (live && sub == 0) {
            const float v = walk_bin_value(a, sum, volume);
            a.out[bin] = a.accumulate ? float(double(a.out[bin]) + double(v)) : v;
            if (MOMENTS) {
                if (a.sum_f)  a.sum_f[bin - a.bin_begin]  = sum;
                if (a.sum_f2) a.sum_f2[bin - a.bin_begin] = sum2;
            }
        }
        signal_tile_done(a.signal, tile, ntiles, lane);
        if (lane == 0) tile = atomicAdd(a.tile_counter, 1ull);
        tile = __shfl_sync(0xffffffffu, tile, 0);
    }
}

// ---- block-fed kernel, one lane per bin, with a two-tile window (the shipped C5 path) -----------------------------------------------
// With one lane per bin (every large grid: pick_lanes_per_bin) a bin is `spp` sequential paths of random length, so the 32 lanes of a
// tile finish at different times and walk_block_kernel idles the early ones until the slowest is through (ncu: 27.5 of 32 lanes active
// at C5).  Here a warp holds TWO tiles: a lane that has stored its bin of the current tile moves straight on to its bin of the next one
// and only waits if it has finished that one too while some lane is still in the current tile (a full bin's lead: it does not happen
// at C5's 256 paths per bin).  When every lane has left the current tile the warp signals it (end-to-end path: chunk flags), the next
// tile becomes the current one and a new ticket is drawn.  Tiles, tickets, elements and the per-bin summation order (one lane, samples
// in order) are walk_block_kernel's at lanes_per_bin = 1: bits identical.
// TAIL: the range has explicit entries beyond the binned dimensions (elements DIMBINS..3 of block 0 are mapped too); without them those
// elements keep [0,1) — fmaf(u, 1, 0) == u — and block 0 costs DIMBINS predicated FFMAs instead of a divergent branch.
#ifndef WALK_WINDOW_MIN_CTAS
#define WALK_WINDOW_MIN_CTAS 5
#endif
template<class F, int DIMBINS, bool MOMENTS, bool EXACT, bool TAIL>
__global__ void __launch_bounds__(MC_THREADS, WALK_WINDOW_MIN_CTAS)
walk_block_window_kernel(const F f, const vb200_walk_launch a) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t nshard = a.bin_end - a.bin_begin;
    const uint64_t ntiles = (nshard + 31u) / 32u;
    // the bin box of this lane's bin in tile t (all lanes together: the index divisions run convergent, once per tile)
    auto prepare = [&](uint64_t t, float (&lo)[DIMBINS], float (&ext)[DIMBINS], float& vol) -> bool {
        const uint64_t b = a.bin_begin + t * 32u + lane;
        if (t >= ntiles || b >= a.bin_end) return false;
        vol = walk_bin_box<DIMBINS>(a.domain, b, lo, ext);
        return true;
    };
    unsigned long long t2 = 0;
    if (lane == 0) t2 = atomicAdd(a.tile_counter, 2ull);          // tickets are consecutive: the first two in one atomic
    t2 = __shfl_sync(0xffffffffu, t2, 0);
    uint64_t tile_cur = t2, tile_nxt = t2 + 1;
    uint32_t base = 0;                    // warp-uniform: tiles this warp has retired
    uint32_t pos = 0;                     // per lane: base = in the current tile, base + 1 = in the next one, base + 2 = through with both (waits)
    // element -> value map of block 0: v = fmaf(u, scale, offset); binned dimensions use the bin box, other explicit entries their
    // range, everything else [0,1) (scale 1, offset 0: fmaf(u,1,0) == u)
    float sc[4] = {1.0f, 1.0f, 1.0f, 1.0f}, of[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (TAIL) {
#pragma unroll
        for (int d = DIMBINS; d < 4; ++d)
            if (d < a.domain.dim) { sc[d] = a.domain.rmax[d] - a.domain.rmin[d]; of[d] = a.domain.rmin[d]; }
    }
    float n_lo[DIMBINS], n_ext[DIMBINS], n_vol = 1.0f, volume = 1.0f;     // n_*: this lane's bin of the NEXT tile, prepared ahead
    float sum = 0.0f, sum2 = 0.0f;
    bool alive = false;
    bool has_bin = prepare(tile_cur, n_lo, n_ext, volume);
    uint64_t bin = a.bin_begin + tile_cur * 32u + lane;
#pragma unroll
    for (int d = 0; d < DIMBINS && d < 4; ++d) { sc[d] = n_ext[d]; of[d] = n_lo[d]; }
    if (!has_bin) pos = 1;
    const uint32_t idle = a.spp + 1u;       // `next` of a lane without a bin: neither < spp (would start a path) nor == spp (bin finished)
    bool n_live = prepare(tile_nxt, n_lo, n_ext, n_vol);
    uint32_t b1 = uint32_t(bin >> 32), next = has_bin ? 0u : idle, blk = 0;
    // first Philox round from cached products (philox4x32_from_products): counter = (bin lo, bin hi, sample, block)
    uint64_t p_bin = uint64_t(uint32_t(bin)) * philox_m0();      // M0 * bin lo: fixed while the lane works on this bin
    const PhiloxKeys<10> keys = philox_key_schedule<10>(a.key0, a.key1);
    uint64_t p_next = 0, p_s = 0;                                 // M1 * next, M1 * (sample in flight): samples count up, so += M1
    typename F::State st;
    while (true) {
        // ---- service: first pass, and whenever some lane is through with its bin -----------------------------------------------------
        if (!alive && next == a.spp) {
            const float v = walk_bin_value(a, sum, volume);
            a.out[bin] = a.accumulate ? float(double(a.out[bin]) + double(v)) : v;
            if (MOMENTS) {
                if (a.sum_f)  a.sum_f[bin - a.bin_begin]  = sum;
                if (a.sum_f2) a.sum_f2[bin - a.bin_begin] = sum2;
            }
            has_bin = false; next = idle; ++pos;
        }
        while (true) {                    // warp-uniform: move free lanes into the next tile; retire the current tile once everybody has left it
            if (!has_bin && pos == base + 1u) {
                if (n_live) {
#pragma unroll
                    for (int d = 0; d < DIMBINS && d < 4; ++d) { sc[d] = n_ext[d]; of[d] = n_lo[d]; }
                    volume = n_vol;
                    bin = a.bin_begin + tile_nxt * 32u + lane; b1 = uint32_t(bin >> 32);
                    p_bin = uint64_t(uint32_t(bin)) * philox_m0(); p_next = 0;
                    sum = 0.0f; sum2 = 0.0f; next = 0; has_bin = true;
                }
                else ++pos;               // no bin for this lane there (ragged last tile, or past the last ticket)
            }
            if (!__all_sync(0xffffffffu, pos > base)) break;
            if (tile_cur >= ntiles) return;                            // tickets only grow: nothing left for this warp
            signal_tile_done(a.signal, tile_cur, ntiles, lane);
            tile_cur = tile_nxt;
            unsigned long long t = 0;
            if (lane == 0) t = atomicAdd(a.tile_counter, 1ull);
            tile_nxt = __shfl_sync(0xffffffffu, t, 0);
            ++base;
            n_live = prepare(tile_nxt, n_lo, n_ext, n_vol);
        }
        // ---- hot loop: one Philox block per lane and iteration, until a lane has finished its bin (some lane always owns a bin here) ----
        do {
            const bool starting = !alive && next < a.spp;
            if (starting) { p_s = p_next; p_next += philox_m1(); ++next; blk = 0; }
            const u32x4 r = philox4x32_from_products<10>(p_bin, p_s, b1, blk, keys);
            BlockIterator it;
            it.e0 = u01(r.x); it.e1 = u01(r.y); it.e2 = u01(r.z); it.e3 = u01(r.w); it.i = 0;
            if (TAIL) {
                if (blk == 0) {       // PhiloxSequence::const_iterator::load: fmaf(u, extent, lower)
                    it.e0 = fmaf(it.e0, sc[0], of[0]); it.e1 = fmaf(it.e1, sc[1], of[1]); it.e2 = fmaf(it.e2, sc[2], of[2]); it.e3 = fmaf(it.e3, sc[3], of[3]);
                }
            } else {
                const bool first = blk == 0;
                if (DIMBINS > 0) it.e0 = first ? fmaf(it.e0, sc[0], of[0]) : it.e0;
                if (DIMBINS > 1) it.e1 = first ? fmaf(it.e1, sc[1], of[1]) : it.e1;
                if (DIMBINS > 2) it.e2 = first ? fmaf(it.e2, sc[2], of[2]) : it.e2;
            }
            ++blk;
            const bool inflight = alive || starting;
            if (starting) {                                              // elements 0,1 — or none, then the first round takes them
                st = f.begin(it); alive = true;
                if constexpr (F::elements_begin == 0) alive = f.step(st, it);
            }
            else if (alive) alive = f.step(st, it);                      // elements 0,1 (or only 0)
            it.i = 2;
            if (alive) alive = f.step(st, it);                           // elements 2,3 (or only 2)
            if (inflight && !alive) {                                    // the path ended in this block
                const float v = f.end(st);
                sum += v;
                if (MOMENTS) sum2 = fmaf(v, v, sum2);
            }
        } while (!__any_sync(0xffffffffu, !alive && next == a.spp));
    }
}

// Replay of recorded sequences (the reference's own element values): one thread per bin, paths in order,
// bins(p) += f(seq)*factor with float(double(acc)+double(f)*factor)  (monte-carlo-per-bin-parallel.h:96).
struct RecordedSequence {
    const float* e; uint32_t len; int32_t* error_flag;
    class const_iterator {
        const RecordedSequence* q; uint32_t i; float n;
        __device__ __forceinline__ void load() {
            // past the recorded end: flag the error and hand out +inf, which ends every `u < p` Russian-roulette test
            // (NaN would make `if (u >= p) break;` spin forever)
            if (i < q->len) n = q->e[i]; else { n = CUDART_INF_F; *q->error_flag = 1; }
        }
    public:
        __device__ __forceinline__ const_iterator(const RecordedSequence* q_) : q(q_), i(0) { load(); }
        __device__ __forceinline__ const float& operator*() const { return n; }
        __device__ __forceinline__ const_iterator& operator++() { ++i; load(); return *this; }
    };
    __device__ __forceinline__ const_iterator begin() const { return const_iterator(this); }
};

template<class F, int DIMBINS, bool EXACT>
__global__ void __launch_bounds__(128)
walk_replay_kernel(const F f, const vb200_walk_replay_launch a) {
    const uint64_t k = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t bin = a.bin_begin + k;
    if (bin >= a.bin_end) return;
    const bool per_bin_mc = a.flavor == VB200_PER_BIN_MC;
    double factor = a.factor;
    if (per_bin_mc) { float lo[DIMBINS], ext[DIMBINS]; factor = __ddiv_rn(double(walk_bin_box<DIMBINS>(a.domain, bin, lo, ext)), double(a.spp)); }
    float acc = per_bin_mc ? 0.0f : a.out[bin];
    for (uint32_t s = 0; s < a.spp; ++s) {
        const uint64_t p = k * a.spp + s;
        RecordedSequence seq{a.elems + a.offsets[p], uint32_t(a.offsets[p + 1] - a.offsets[p]), a.error_flag};
        const float v = f(seq);
        acc = __double2float_rn(__dadd_rn(double(acc), __dmul_rn(double(v), factor)));
    }
    a.out[bin] = per_bin_mc ? __double2float_rn(__dmul_rn(double(a.nbins_total), double(acc))) : acc;
}

// K4 over an infinite range — reference MonteCarlo::integrate(RangeInfinite), src/monte-carlo/monte-carlo.h:65-84: every sample is
// its own lazy sequence (upstream: a fresh mt19937 per sample; here Philox keyed by the global sample index), the bin comes from
// its first DIMBINS elements (:76-80) and bins(pos) += f(seq)*factor.  Same privatised-histogram scheme as mc_scatter_kernel.
struct GlobalSequence {
    uint32_t s0, s1, k0, k1; const vb200_domain* dom;
    class const_iterator {
        const GlobalSequence* q; uint32_t i; u32x4 blk; float n;
        __device__ __forceinline__ void load() {
            if ((i & 3u) == 0u) blk = philox4x32<10>(u32x4{q->s0, q->s1, 0xfffffffeu, i >> 2}, q->k0, q->k1);
            const float u = pick(blk, int(i & 3u));
            n = int(i) < q->dom->dim ? fmaf(u, q->dom->rmax[i] - q->dom->rmin[i], q->dom->rmin[i]) : u;
        }
    public:
        __device__ __forceinline__ const_iterator(const GlobalSequence* q_) : q(q_), i(0) { load(); }
        __device__ __forceinline__ const float& operator*() const { return n; }
        __device__ __forceinline__ const_iterator& operator++() { ++i; load(); return *this; }
    };
    __device__ __forceinline__ const_iterator begin() const { return const_iterator(this); }
};

template<class F, int DIMBINS, bool EXACT>
__global__ void __launch_bounds__(256)
walk_scatter_kernel(const F f, const vb200_scatter_launch a) {
    constexpr int SMEM_BINS = 8192;
    __shared__ float s_hist[SMEM_BINS];
    const bool priv = a.nbins_total <= uint64_t(SMEM_BINS);
    if (priv) { for (uint32_t i = threadIdx.x; i < a.nbins_total; i += blockDim.x) s_hist[i] = 0.0f; __syncthreads(); }
    const float factor = float(a.factor);
    for (uint64_t s = a.sample_begin + uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; s < a.sample_end; s += uint64_t(gridDim.x) * blockDim.x) {
        GlobalSequence seq{uint32_t(s), uint32_t(s >> 32), a.key0, a.key1, &a.domain};
        uint64_t lin = 0, prod = 1;
        {
            auto it = seq.begin();
#pragma unroll
            for (int i = 0; i < DIMBINS; ++i) {
                const float lo = i < a.domain.dim ? a.domain.rmin[i] : 0.0f, hi = i < a.domain.dim ? a.domain.rmax[i] : 1.0f;
                const float t = float(a.domain.res[i]) * (*it - lo) / (hi - lo);          // monte-carlo.h:79
                uint64_t p = uint64_t(t); if (p >= a.domain.res[i]) p = a.domain.res[i] - 1;
                lin += p * prod; prod *= a.domain.res[i];
                if (i + 1 < DIMBINS) ++it;
            }
        }
        const float v = f(seq) * factor;
        if (priv) atomicAdd(&s_hist[lin], v); else atomicAdd(&a.out[lin], v);
    }
    if (priv) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < a.nbins_total; i += blockDim.x) { const float v = s_hist[i]; if (v != 0.0f) atomicAdd(&a.out[i], v); }
    }
}

}}} // namespace viltrum::b200::device
