// Internal (library-private) context: device, stream, scratch arena, error text.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <cstdio>
#include <cstdarg>
#include <atomic>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <vector>
#include <utility>
#include "../../include/viltrum_b200.h"

namespace vb200 {
// Host side of the end-to-end sampling path: the kernel raises one host-mapped flag per chunk of bins; whoever claims a chunk
// spins on its flag and then applies the reference's '+=' / '=' to the caller's bins.  One thread cannot keep up with the
// kernel (a 4 MiB '+=' pass takes longer than the 0.34 ms the B200 needs to produce it), so a few pooled workers share the chunks.
struct HostPass {
    volatile uint32_t* flags = nullptr; uint32_t epoch = 0;
    uint64_t chunks = 0, bins_per_chunk = 0, n = 0;
    float* dst = nullptr; const float* src = nullptr; bool accumulate = false;
    std::atomic<uint64_t> next{0}, finished{0};
    std::atomic<int> abort{0};
    void run(bool poll_stream, cudaStream_t stream, cudaError_t* stream_error);
};
struct HostPool {
    std::vector<std::thread> workers;
    std::mutex m; std::condition_variable cv;
    HostPass* job = nullptr; uint64_t job_id = 0; bool stop = false;
    std::atomic<int> active{0};
    void start(int nworkers);
    void publish(HostPass* j);
    void retire();              // no worker touches the job after this returns
    ~HostPool();
};
}

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
    // sampling kernels: d_counter[0] = dynamic tile scheduler ticket; the kMaxChunks uint32 words behind it (d_done) count the
    // finished tiles per chunk of the end-to-end path; h_flags (pinned, device-mapped) receive the per-chunk completion epochs
    unsigned long long* d_counter = nullptr;
    uint32_t* d_done = nullptr;
    uint32_t* h_flags = nullptr;
    uint32_t epoch = 0;
    bool counters_dirty = true;     // d_counter / d_done need a memset before the next sampling launch
    static constexpr int kMaxChunks = 1024;
    vb200::HostPool pool;
    // caller-owned host buffers pinned and mapped by vb200_host_register: (base, bytes)
    std::vector<std::pair<char*, size_t>> registered;
    // region tables handed out and not yet freed: vb200_destroy releases their device memory and orphans them, so that a late
    // vb200_regions_free (a Python finaliser running after Context.close) only deletes the host record
    std::vector<struct vb200_regions*> live_regions;
    // optional NCCL communicator (vb200_comm_init; comm.cu): one rank per process and GPU
    void* comm = nullptr; int comm_rank = 0, comm_size = 1;
    // vb200_kernel_timer: event pairs around the launches of the residual-sampling kernel
    bool ktimer = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ktimer_events;
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
uint32_t pick_lanes_per_bin(const vb200_ctx* ctx, uint64_t spp, uint64_t nbins, uint32_t group);
// bins staging helpers: returns the device pointer to use as "base of the full grid"
struct BinStage { float* dev_base = nullptr; bool staged = false; uint64_t begin = 0, end = 0; float* host = nullptr; };
int stage_bins_in(vb200_ctx* ctx, float* bins, int mem, uint64_t begin, uint64_t end, bool upload, BinStage* st);
int stage_bins_out(vb200_ctx* ctx, const BinStage& st);
void comm_release(vb200_ctx* ctx);
int comm_allreduce_sum(vb200_ctx* ctx, float* dev, uint64_t count);
int add_into(vb200_ctx* ctx, float* bins, const float* add, uint64_t n);

} // namespace vb200
