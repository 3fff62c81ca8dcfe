// Experiment harness (not product): wavefront walk kernel variants (refill threshold, steps per iteration) on the C5 workload.
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include <viltrum_b200/device/walk.cuh>
#include "../../viltrum_b200/csrc/builtin_integrands.cuh"
using namespace viltrum::b200;

// EXPERIMENT (measured slower than the shipped wavefront kernel: profiles/walk_variants_r1b.txt — 39.4 vs 47.4 G paths/s at LPB 1)
namespace viltrum { namespace b200 { namespace device {
// ---- wavefront with a prefetched generator block: per-lane refill at no extra generator cost ------------------------------------
// The generator is the expensive part of a path (20 IMAD.WIDE per Philox call, profiles/pipes_r1.txt), so it must run ONCE per
// loop iteration for the whole warp.  In walk_wavefront_kernel a re-armed lane pays a second, divergent call for block 0 of its
// new sample, which is why dead lanes were only re-armed in batches of 8.  Here every iteration starts with one convergent call
// in which a live lane generates the NEXT block of its path and a dead lane block 0 of its NEXT sample; the iterator swaps the
// prefetched block in when the path crosses into it (and falls back to generating on demand if a functor consumes its elements
// in another rhythm, so the stream element i = Philox(bin, sample, i/4)[i%4] holds for any functor).  With two roulette rounds
// per iteration a live lane consumes exactly one block per iteration — round 1 takes elements 2,3, round 2 elements 0,1 of the
// prefetched block — and a fresh path's begin() (elements 0,1 of block 0) runs in the slot of round 2, so all lanes stay
// aligned and a dead lane is back at work within one iteration.  Same elements, same per-lane sample order: bins are
// bit-identical to walk_kernel's.
template<int DIMBINS>
struct PrefetchedSequenceIterator {
    const PhiloxSequence<DIMBINS>* q; uint32_t i; u32x4 blk, nxt; uint32_t nxt_block; float n;
    __device__ __forceinline__ void load() {
        if ((i & 3u) == 0u) {
            const uint32_t b = i >> 2;
            blk = (b == nxt_block) ? nxt : philox4x32<10>(u32x4{q->b0, q->b1, q->s, b}, q->k0, q->k1);
        }
        const float u = pick(blk, int(i & 3u));
        float v = u;
        bool binned = false;
#pragma unroll
        for (int d = 0; d < DIMBINS; ++d) if (i == uint32_t(d)) { v = fmaf(u, q->ext[d], q->lo[d]); binned = true; }
        if (!binned && int(i) < q->dom->dim) v = fmaf(u, q->dom->rmax[i] - q->dom->rmin[i], q->dom->rmin[i]);
        n = v;
    }
    __device__ __forceinline__ explicit PrefetchedSequenceIterator(const PhiloxSequence<DIMBINS>* q_) : q(q_), i(0), blk{0, 0, 0, 0}, nxt{0, 0, 0, 0}, nxt_block(0xffffffffu), n(0.0f) {}
    __device__ __forceinline__ void restart() { i = 0; load(); }          // block 0 of q->s must have been prefetched (or is generated here)
    __device__ __forceinline__ const float& operator*() const { return n; }
    __device__ __forceinline__ PrefetchedSequenceIterator& operator++() { ++i; load(); return *this; }
};

template<class F, int DIMBINS, bool MOMENTS, bool EXACT>
__global__ void __launch_bounds__(MC_THREADS)
walk_prefetch_kernel(const F f, const vb200_walk_launch a) {
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
        PrefetchedSequenceIterator<DIMBINS> it(&seq);
        while (true) {
            const bool want = !alive && next < a.spp;
            if (!__any_sync(0xffffffffu, alive || want)) break;
            // 1. ONE generator call for the warp: the block a live path is about to cross into, or block 0 of a fresh sample
            if (alive || want) {
                const uint32_t smp = alive ? seq.s : next;
                const uint32_t blk = alive ? (it.i >> 2) + 1u : 0u;
                it.nxt = philox4x32<10>(u32x4{seq.b0, seq.b1, smp, blk}, seq.k0, seq.k1);
                it.nxt_block = blk;
            }
            // 2. first roulette round of the live lanes
            const bool was_alive = alive;
            if (alive && !f.step(st, it)) {
                const float v = f.end(st);
                sum += v; if (MOMENTS) sum2 = fmaf(v, v, sum2);
                alive = false;
            }
            // 3. fresh paths start in the slot of the second round ...
            if (want) {
                seq.s = next; next += LPB;
                it.restart();
                st = f.begin(it);
                alive = true;
            } else if (was_alive && alive && !f.step(st, it)) {       // ... which the surviving lanes use for their second round
                const float v = f.end(st);
                sum += v; if (MOMENTS) sum2 = fmaf(v, v, sum2);
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

}}}
template<class K> int occ_grid(K k, int sms) { int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 256, 0); return occ * sms; }
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    vb200_walk_launch L; memset(&L, 0, sizeof(L));
    L.domain.dim = 0; L.domain.dimbins = 2; L.domain.res[0] = 2048; L.domain.res[1] = 2048; L.domain.drange[0] = L.domain.drange[1] = 1.0f / 2048.0f;
    const uint64_t nb = 1ull << 22; L.bin_begin = 0; L.bin_end = nb; L.nbins_total = nb; L.spp = 256; L.key0 = 1; L.key1 = 2; L.factor = 1.0 / 256;
    cudaMalloc(&L.out, nb * 4); unsigned long long* ctr; cudaMalloc(&ctr, 8); L.tile_counter = ctr;
    builtin::Walk f; builtin::WalkPlain fp;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto k, auto fun, uint32_t lpb) {
        L.lanes_per_bin = lpb; int g = occ_grid(k, sms); float best = 1e30f;
        for (int i = 0; i < 4; ++i) { cudaMemsetAsync(ctr, 0, 8); cudaEventRecord(e0); k<<<g, 256>>>(fun, L); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (i) best = fminf(best, ms); }
        printf("%-40s LPB %2u grid %5d  %7.2f ms  %6.1f G paths/s  (%s)\n", name, lpb, g, best, double(nb) * 256 / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
    };
    for (uint32_t lpb : {1u, 4u, 8u, 32u}) {
        run("generic per-lane", device::walk_kernel<builtin::WalkPlain, 2, false, false>, fp, lpb);
        run("wavefront, prefetched block, per-lane refill", device::walk_prefetch_kernel<builtin::Walk, 2, false, false>, f, lpb);
        run("wavefront refill 8  steps 1", device::walk_wavefront_kernel<builtin::Walk, 2, false, false, 8, 1>, f, lpb);
        run("wavefront refill 12 steps 1", device::walk_wavefront_kernel<builtin::Walk, 2, false, false, 12, 1>, f, lpb);
        run("wavefront refill 8  steps 2", device::walk_wavefront_kernel<builtin::Walk, 2, false, false, 8, 2>, f, lpb);
        run("wavefront refill 12 steps 2", device::walk_wavefront_kernel<builtin::Walk, 2, false, false, 12, 2>, f, lpb);
        run("wavefront refill 16 steps 2", device::walk_wavefront_kernel<builtin::Walk, 2, false, false, 16, 2>, f, lpb);
        run("wavefront refill 24 steps 2", device::walk_wavefront_kernel<builtin::Walk, 2, false, false, 24, 2>, f, lpb);
        run("wavefront refill 16 steps 4", device::walk_wavefront_kernel<builtin::Walk, 2, false, false, 16, 4>, f, lpb);
    }
    return 0;
}
