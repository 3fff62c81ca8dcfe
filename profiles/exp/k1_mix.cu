// Experiment harness (not product): the PRODUCT's mc_per_bin_kernel (include/viltrum_b200/device/mc_per_bin.cuh) on the C2 workload
// (1024x1024 bins, 64 spp, shade4<64>), one executable per generator mix: build with
//   -DVB200_MC_TF_NUM=n (n of 5 generator calls per 8-sample group from Threefry4x32, the rest from Philox4x32-10)
//   -DVB200_MC_TF_ROUNDS=12|20  -DVB200_MC_MINB=k  -DK1_RNG=1 (xoshiro128++ stream per bin, seeded by Philox)
// Prints one line: variant, registers, ms, G evals/s, fraction of the nominal FP32 peak, mean of the bins.
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include <viltrum_b200/device/mc_per_bin.cuh>
#include "../../viltrum_b200/csrc/builtin_integrands.cuh"
using namespace viltrum::b200;
#ifndef VARIANT
#define VARIANT "?"
#endif
int main(int argc, char** argv) {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int reps = argc > 1 ? atoi(argv[1]) : 20;
    vb200_mc_launch a; std::memset(&a, 0, sizeof(a));
    a.domain.dim = 4; a.domain.dimbins = 2;
    for (int i = 0; i < 4; ++i) { a.domain.rmin[i] = 0.f; a.domain.rmax[i] = 1.f; }
    a.domain.res[0] = a.domain.res[1] = 1024; a.domain.drange[0] = a.domain.drange[1] = 1.0f / 1024.0f;
    a.bin_begin = 0; a.bin_end = a.nbins_total = 1u << 20; a.spp = 64; a.lanes_per_bin = 1; a.key0 = 1; a.key1 = 2;
    a.flavor = VB200_MC_PER_BIN; a.accumulate = 0; a.factor = 1.0 / 64.0; a.narrow_binned = 1;
    cudaMalloc(&a.out, a.nbins_total * 4); cudaMalloc(&a.tile_counter, 16); cudaMemset(a.tile_counter, 0, 16);
    #ifndef K1_RNG
#define K1_RNG 1      // device::MC_RNG_PHILOX; 0 = MC_RNG_XOSHIRO
#endif
    auto k = device::mc_per_bin_kernel<builtin::Shade4<64>, 4, 2, false, false, true, K1_RNG>;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, device::MC_THREADS, 0);
    const int grid = occ * sms;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto launch = [&] { k<<<grid, device::MC_THREADS>>>(builtin::Shade4<64>(), a); };      // the kernel resets its scheduler words itself
    for (int i = 0; i < 5; ++i) launch();
    cudaDeviceSynchronize();
    float best = 1e30f, tot = 0;
    for (int i = 0; i < reps; ++i) {
        cudaEventRecord(e0); k<<<grid, device::MC_THREADS>>>(builtin::Shade4<64>(), a); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); tot += ms; if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
    std::vector<float> h(a.nbins_total); cudaMemcpy(h.data(), a.out, h.size() * 4, cudaMemcpyDeviceToHost);
    double m = 0, m2 = 0; for (float v : h) { m += v; m2 += double(v) * v; } m /= h.size(); m2 /= h.size();
    const double evals = double(a.nbins_total) * a.spp, ms = tot / reps;
    printf("%-28s regs %3d occ %d grid %4d  mean %.4f ms (best %.4f)  %6.1f G evals/s  frac %.4f  bins mean %.6f var %.6f\n", VARIANT, fa.numRegs, occ, grid, ms, best,
           evals / ms * 1e-6, evals * 155 / ms * 1e-9 / (2.0 * 128 * sms * 1.965e-3), m, m2 - m * m);
    return 0;
}
