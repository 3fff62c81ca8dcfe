// Internal (library-private) context: device, stream, scratch arena, error text.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <cstdio>
#include <cstdarg>
#include "../../include/viltrum_b200.h"

struct vb200_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::string error;
    uint64_t launches = 0;
    // grow-only device scratch for staging VB200_HOST arguments
    void* scratch[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t scratch_bytes[4] = {0, 0, 0, 0};
    // pinned host staging (D2H of bins lands here first so the copy is truly asynchronous)
    void* pinned = nullptr; size_t pinned_bytes = 0;
    int32_t* d_flag = nullptr;      // device error flag for replay kernels
    unsigned long long* d_counter = nullptr;   // dynamic tile schedulers of the sampling kernels: one ticket counter per chunk
    static constexpr int kMaxChunks = 8;
    cudaEvent_t chunk_done[kMaxChunks] = {};
};

namespace vb200 {

int fail(vb200_ctx* ctx, int status, const char* fmt, ...);
#define VB200_CUDA(ctx, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) \
    return vb200::fail(ctx, VB200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); } while (0)

int reserve(vb200_ctx* ctx, int slot, size_t bytes, void** out);
// Stream-ordered device memory from the device's default CUDA memory pool (cudaMallocAsync on the context's stream; the pool's
// release threshold is raised at context creation so freed blocks stay cached): the region pipelines allocate and free
// hundreds of MB per call, and plain cudaMalloc/cudaFree (each an implicit device synchronisation) dominated their wall time.
cudaError_t dmalloc_bytes(vb200_ctx* ctx, void** p, size_t bytes);
template<class T> inline cudaError_t dmalloc(vb200_ctx* ctx, T** p, size_t bytes) { return dmalloc_bytes(ctx, reinterpret_cast<void**>(p), bytes); }
void dfree(vb200_ctx* ctx, void* p);
int reserve_pinned(vb200_ctx* ctx, size_t bytes, void** out);
inline uint64_t nbins_of(const vb200_domain& d) { uint64_t n = 1; for (int i = 0; i < d.dimbins; ++i) n *= d.res[i]; return n; }
int check_domain(vb200_ctx* ctx, const vb200_domain& d, int integrand_dim);
// copy of the caller's domain with drange filled in: (max-min)/float(res) in fp32, implicit [0,1] beyond d.dim (infinite ranges)
vb200_domain finish_domain(const vb200_domain& d);
int resolve_shard(vb200_ctx* ctx, const vb200_shard& s, uint64_t total, uint64_t* begin, uint64_t* end);
int call_thunk(vb200_ctx* ctx, const vb200_integrand* f, int kind, const void* args);
uint32_t pick_lanes_per_bin(const vb200_ctx* ctx, uint64_t spp, uint64_t nbins);
// bins staging helpers: returns the device pointer to use as "base of the full grid"
struct BinStage { float* dev_base = nullptr; bool staged = false; uint64_t begin = 0, end = 0; float* host = nullptr; };
int stage_bins_in(vb200_ctx* ctx, float* bins, int mem, uint64_t begin, uint64_t end, bool upload, BinStage* st);
int stage_bins_out(vb200_ctx* ctx, const BinStage& st);

} // namespace vb200
