// Experiment harness (not product): wavefront walk kernel variants (refill threshold, steps per iteration) on the C5 workload.
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include <viltrum_b200/device/walk.cuh>
#include "../../viltrum_b200/csrc/builtin_integrands.cuh"
using namespace viltrum::b200;
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
