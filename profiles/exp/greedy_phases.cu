// Experiment harness (not product): the PRODUCT's greedy_kernel (include/viltrum_b200/device/greedy.cuh) on BASELINE config 3's shape
// (smooth_edge2, nested(boole,simpson), size/relative 1e-5) with per-phase cycle counters of thread 0 (-DVB200_GREEDY_TIMING).
//   phases (thread 0's clock): 0 barrier B1 | 1 pop | 2 wait at barrier B2 (the workers' fetch + split + errors + picks, if longer than the pop) |
//           3 two pushes + next top.  (Round-2 first capture, old kernel: top 302, pop 5223, split 970, errors 1378, pushes+stores 2001 of 10234 cycles.)
#define VB200_GREEDY_TIMING
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cuda_runtime.h>
#include <viltrum_b200/device/greedy.cuh>
#include "../../viltrum_b200/csrc/builtin_integrands.cuh"
using namespace viltrum::b200;
int main(int argc, char** argv) {
    const uint64_t it = argc > 1 ? strtoull(argv[1], 0, 10) : 200000;
    vb200_greedy_launch a; std::memset(&a, 0, sizeof(a));
    const int D = 2, SD = 25;
    a.dim = D; a.rule = VB200_RULE_BOOLE_SIMPSON; a.heuristic = VB200_HEURISTIC_SIZE; a.metric = VB200_METRIC_RELATIVE; a.size_weight = 1e-5;
    a.iterations = it; a.capacity = 2 * it + 1;
    cudaMalloc(&a.range, a.capacity * 2 * D * 4); cudaMalloc(&a.data, a.capacity * SD * 4); cudaMalloc(&a.err, a.capacity * 4);
    cudaMalloc(&a.heap, (it + 2) * 8); cudaMalloc(&a.heap_size, 8);
    a.range_min[0] = a.range_min[1] = 0.f; a.range_max[0] = a.range_max[1] = 1.f;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        unsigned long long zero[8] = {0}; cudaMemcpyToSymbol(device::vb200_greedy_clock, zero, sizeof(zero));
        cudaEventRecord(e0);
        int rc = device::launch_greedy<builtin::SmoothEdge2, 2, true, float>(builtin::SmoothEdge2(), a, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        unsigned long long c[8]; cudaMemcpyFromSymbol(c, device::vb200_greedy_clock, sizeof(c));
        printf("rc %d: %llu iterations in %.1f ms = %.3f us/iteration; cycles per iteration by phase:", rc, (unsigned long long)it, ms, ms * 1e3 / it);
        unsigned long long tot = 0; for (int k = 0; k < 7; ++k) { printf(" %d:%.0f", k, double(c[k]) / it); tot += c[k]; }
        printf("  total %.0f (%s)\n", double(tot) / it, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
