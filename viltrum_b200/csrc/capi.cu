// libviltrum_b200.so — context, argument staging and the Monte-Carlo drivers of the C ABI (include/viltrum_b200.h).
// Nothing in here evaluates an integrand: kernels templated on the functor are reached through the launch thunks
// in the vb200_integrand table.  There is no CPU fallback anywhere in this library.
#include "context.h"
#include "regions.h"
#include <chrono>
#include <viltrum_b200/device/philox.cuh>
#include <viltrum_b200/device/xoshiro.cuh>
#include <viltrum_b200/device/threefry.cuh>
#include <cstring>
#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <new>
#include <atomic>

namespace {
thread_local std::string g_create_error;
}

namespace vb200 {

int fail(vb200_ctx* ctx, int status, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    if (ctx) { ctx->error = buf; ctx->counters_dirty = true; } else g_create_error = buf;      // after any failure the sampler's scheduler words are cleared again
    return status;
}

void HostPass::run(bool poll_stream, cudaStream_t stream, cudaError_t* stream_error) {
    for (;;) {
        const uint64_t c = next.fetch_add(1, std::memory_order_relaxed);
        if (c >= chunks) return;
        uint64_t spins = 0;
        while (flags[c] != epoch) {
            if (abort.load(std::memory_order_relaxed)) return;
            if ((++spins & 0x3ffu) == 0) std::this_thread::yield();      // be polite if the cores are oversubscribed
            if (poll_stream && (spins & 0x3fffu) == 0) {       // the kernel may have died: do not spin forever
                const cudaError_t q = cudaStreamQuery(stream);
                if (q != cudaErrorNotReady && flags[c] != epoch) { *stream_error = (q == cudaSuccess) ? cudaErrorUnknown : q; abort.store(1); return; }
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        const uint64_t lo = c * bins_per_chunk, hi = (lo + bins_per_chunk < n) ? lo + bins_per_chunk : n;
        // float(double(dst)+double(src)) == dst+src in fp32 (the double sum of two floats rounds to the same float)
        if (accumulate) { float* __restrict__ d = dst; const float* __restrict__ s = src; for (uint64_t i = lo; i < hi; ++i) d[i] += s[i]; }
        else std::memcpy(dst + lo, src + lo, (hi - lo) * sizeof(float));
        finished.fetch_add(1, std::memory_order_release);
    }
}

void HostPool::start(int nworkers) {
    while (int(workers.size()) < nworkers) {
        workers.emplace_back([this] {
            uint64_t seen = 0;
            for (;;) {
                HostPass* j = nullptr;
                {
                    std::unique_lock<std::mutex> lk(m);
                    cv.wait(lk, [&] { return stop || (job && job_id != seen); });
                    if (stop) return;
                    seen = job_id; j = job; active.fetch_add(1);
                }
                cudaError_t unused = cudaSuccess;
                j->run(false, nullptr, &unused);
                active.fetch_sub(1, std::memory_order_release);
            }
        });
    }
}
void HostPool::publish(HostPass* j) {
    { std::lock_guard<std::mutex> lk(m); job = j; ++job_id; }
    cv.notify_all();
}
void HostPool::retire() {
    { std::lock_guard<std::mutex> lk(m); job = nullptr; }
    while (active.load(std::memory_order_acquire) != 0) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
}
HostPool::~HostPool() {
    { std::lock_guard<std::mutex> lk(m); stop = true; }
    cv.notify_all();
    for (auto& t : workers) t.join();
}

int reserve(vb200_ctx* ctx, int slot, size_t bytes, void** out) {
    if (bytes > ctx->scratch_bytes[slot]) {
        if (ctx->scratch[slot]) { VB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); VB200_CUDA(ctx, cudaFree(ctx->scratch[slot])); ctx->scratch[slot] = nullptr; ctx->scratch_bytes[slot] = 0; }
        size_t want = bytes + bytes / 4 + 256;
        if (cudaMalloc(&ctx->scratch[slot], want) != cudaSuccess) { cudaGetLastError(); return fail(ctx, VB200_ERR_NOMEM, "cudaMalloc of %zu scratch bytes failed", want); }
        ctx->scratch_bytes[slot] = want;
    }
    *out = ctx->scratch[slot];
    return VB200_OK;
}

cudaError_t dmalloc_bytes(vb200_ctx* ctx, void** p, size_t bytes) {
    *p = nullptr;
    return cudaMallocAsync(p, bytes ? bytes : 1, ctx->stream);
}
void dfree(vb200_ctx* ctx, void* p) { if (p) cudaFreeAsync(p, ctx->stream); }

int reserve_pinned(vb200_ctx* ctx, size_t bytes, void** out) {
    if (bytes > ctx->pinned_bytes) {
        if (ctx->pinned) { VB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); VB200_CUDA(ctx, cudaFreeHost(ctx->pinned)); ctx->pinned = nullptr; ctx->pinned_bytes = 0; }
        size_t want = bytes + bytes / 4 + 256;
        if (cudaMallocHost(&ctx->pinned, want) != cudaSuccess) { cudaGetLastError(); return fail(ctx, VB200_ERR_NOMEM, "cudaMallocHost of %zu bytes failed", want); }
        ctx->pinned_bytes = want;
    }
    *out = ctx->pinned;
    return VB200_OK;
}

int check_domain(vb200_ctx* ctx, const vb200_domain& d, int integrand_dim) {
    if (d.dimbins < 1 || d.dimbins > VB200_MAX_DIMBINS) return fail(ctx, VB200_ERR_INVALID, "dimbins=%d outside 1..%d", d.dimbins, VB200_MAX_DIMBINS);
    if (integrand_dim > 0) {
        if (d.dim != integrand_dim) return fail(ctx, VB200_ERR_INVALID, "range has %d dimensions, integrand takes %d", d.dim, integrand_dim);
        if (d.dimbins > d.dim) return fail(ctx, VB200_ERR_INVALID, "dimbins=%d exceeds range dimensions %d", d.dimbins, d.dim);
    } else if (d.dim < 0 || d.dim > VB200_MAX_DIM) return fail(ctx, VB200_ERR_INVALID, "infinite range with %d explicit entries (max %d)", d.dim, VB200_MAX_DIM);
    for (int i = 0; i < d.dimbins; ++i) if (d.res[i] == 0 || d.res[i] > 0xffffffffull) return fail(ctx, VB200_ERR_INVALID, "resolution[%d]=%llu invalid", i, (unsigned long long)d.res[i]);
    return VB200_OK;
}

int resolve_shard(vb200_ctx* ctx, const vb200_shard& s, uint64_t total, uint64_t* begin, uint64_t* end) {
    if (s.begin == 0 && s.end == 0) { *begin = 0; *end = total; return VB200_OK; }
    if (s.begin == VB200_SHARD_EMPTY_INDEX && s.end == VB200_SHARD_EMPTY_INDEX) { *begin = 0; *end = 0; return VB200_OK; }      // explicit empty shard
    if (s.begin > s.end || s.end > total) return fail(ctx, VB200_ERR_INVALID, "shard [%llu,%llu) outside [0,%llu)", (unsigned long long)s.begin, (unsigned long long)s.end, (unsigned long long)total);
    *begin = s.begin; *end = s.end;
    return VB200_OK;
}

int call_thunk(vb200_ctx* ctx, const vb200_integrand* f, int kind, const void* args) {
    if (!f || f->abi_version != VB200_ABI_VERSION) return fail(ctx, VB200_ERR_INVALID, "integrand descriptor missing or built against another ABI version");
    if (!f->launch[kind]) return fail(ctx, VB200_ERR_UNSUPPORTED, "integrand '%s' has no kernel of kind %d", f->name ? f->name : "?", kind);
    int e = f->launch[kind](f, args, ctx->stream);
    if (e != 0) return fail(ctx, VB200_ERR_CUDA, "kernel launch (kind %d, integrand '%s') failed: %s", kind, f->name ? f->name : "?", cudaGetErrorString(cudaError_t(e)));
    ctx->launches++;
    return VB200_OK;
}

vb200_domain finish_domain(const vb200_domain& d) {
    vb200_domain o = d;
    for (int i = 0; i < VB200_MAX_DIMBINS; ++i) o.drange[i] = 0.0f;
    for (int i = 0; i < d.dimbins && i < VB200_MAX_DIMBINS; ++i) {
        const float lo = i < d.dim ? d.rmin[i] : 0.0f, hi = i < d.dim ? d.rmax[i] : 1.0f;
        o.drange[i] = (hi - lo) / float(d.res[i]);
    }
    return o;
}

// Lanes of one warp that share a bin.  Every bin costs a fixed set-up (bin box, scaling, tile ticket) worth about one
// sample, so as few lanes per bin as possible — but enough lanes in total to fill the chip twice over when the grid is
// small (measured: profiles/mc_variants_r1.txt).
uint32_t pick_lanes_per_bin(const vb200_ctx* ctx, uint64_t spp, uint64_t nbins, uint32_t group) {
    const uint64_t want_lanes = uint64_t(ctx->sm_count) * 2048ull * 2ull;
    const uint64_t groups = (spp + group - 1) / group;      // the lanes of a bin stride over draw groups of eight samples (mc_per_bin.cuh); paths in walk.cuh stride over samples (group = 4 bounds their lanes)
    uint32_t lpb = 1;
    // VB200_LANES_PER_BIN (tuning / test knob): upper bound on the lanes that share a bin; 1 gives small grids the summation order —
    // and the kernels — of the large ones
    uint32_t cap = 32;
    if (const char* env = std::getenv("VB200_LANES_PER_BIN")) { const long v = std::atol(env); if (v >= 1 && v <= 32) cap = uint32_t(v); }
    while (lpb < 32 && lpb * 2 <= cap && nbins * lpb < want_lanes && uint64_t(lpb) * 2 <= groups) lpb <<= 1;
    return lpb;
}

int stage_bins_in(vb200_ctx* ctx, float* bins, int mem, uint64_t begin, uint64_t end, bool upload, BinStage* st) {
    st->begin = begin; st->end = end; st->host = nullptr; st->staged = false;
    if (!bins) return fail(ctx, VB200_ERR_INVALID, "bins pointer is NULL");
    if (mem == VB200_DEVICE) { st->dev_base = bins; return VB200_OK; }
    if (mem != VB200_HOST) return fail(ctx, VB200_ERR_INVALID, "bad memory-space flag %d", mem);
    const uint64_t n = end - begin;
    void* d = nullptr;
    int rc = reserve(ctx, 0, n * sizeof(float), &d); if (rc) return rc;
    if (upload) VB200_CUDA(ctx, cudaMemcpyAsync(d, bins + begin, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    st->dev_base = static_cast<float*>(d) - begin;      // kernels index from the base of the full grid
    st->staged = true; st->host = bins;
    return VB200_OK;
}

int stage_bins_out(vb200_ctx* ctx, const BinStage& st) {
    if (!st.staged) return VB200_OK;
    const uint64_t n = st.end - st.begin;
    VB200_CUDA(ctx, cudaMemcpyAsync(st.host + st.begin, st.dev_base + st.begin, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    VB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VB200_OK;
}

} // namespace vb200

using namespace vb200;

// ------------------------------------------------------------------------------------------------------------
extern "C" int vb200_create(int device, vb200_ctx** out) {
    if (!out) return fail(nullptr, VB200_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { cudaGetLastError(); return fail(nullptr, VB200_ERR_NO_DEVICE, "no CUDA device available (%s); viltrum_b200 has no CPU fallback",
            e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"); }
    if (device < 0 || device >= n) return fail(nullptr, VB200_ERR_INVALID, "device %d outside 0..%d", device, n - 1);
    vb200_ctx* ctx = new (std::nothrow) vb200_ctx;
    if (!ctx) return fail(nullptr, VB200_ERR_NOMEM, "out of host memory");
    ctx->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess ||
        (e = cudaMalloc(&ctx->d_flag, sizeof(int32_t))) != cudaSuccess ||
        (e = cudaMalloc(&ctx->d_counter, 2 * sizeof(unsigned long long) + vb200_ctx::kMaxChunks * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaHostAlloc(&ctx->h_flags, vb200_ctx::kMaxChunks * sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable)) != cudaSuccess) {
        int rc = fail(nullptr, VB200_ERR_CUDA, "context setup on device %d failed: %s", device, cudaGetErrorString(e));
        delete ctx; return rc;
    }
    {   // keep freed blocks in the device's default memory pool instead of returning them to the OS at every synchronisation
        cudaMemPool_t pool = nullptr; unsigned long long keep = ~0ull;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        cudaGetLastError();
    }
    ctx->d_done = reinterpret_cast<uint32_t*>(ctx->d_counter + 2);
    std::memset(ctx->h_flags, 0, vb200_ctx::kMaxChunks * sizeof(uint32_t));
    *out = ctx;
    return VB200_OK;
}

static void ktimer_drop(vb200_ctx* ctx) {
    for (auto& e : ctx->ktimer_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    ctx->ktimer_events.clear();
}
extern "C" void vb200_destroy(vb200_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    ktimer_drop(ctx);               // a timer left on: its events go with the context
    vb200::orphan_regions(ctx);
    vb200::comm_release(ctx);
    for (auto& s : ctx->scratch) if (s) cudaFree(s);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->d_flag) cudaFree(ctx->d_flag);
    if (ctx->d_counter) cudaFree(ctx->d_counter);
    if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
    for (auto& r : ctx->registered) cudaHostUnregister(r.first);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* vb200_last_error(const vb200_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }
extern "C" void* vb200_stream(vb200_ctx* ctx) { return ctx ? ctx->stream : nullptr; }
extern "C" int vb200_synchronize(vb200_ctx* ctx) { if (!ctx) return VB200_ERR_INVALID; VB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); return VB200_OK; }
extern "C" int vb200_sm_count(const vb200_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
extern "C" uint64_t vb200_launch_count(const vb200_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int vb200_kernel_timer(vb200_ctx* ctx, int enable) {
    if (!ctx) return VB200_ERR_INVALID;
    ctx->ktimer = enable != 0;
    if (!ctx->ktimer) { cudaStreamSynchronize(ctx->stream); ktimer_drop(ctx); }
    return VB200_OK;
}
extern "C" int vb200_kernel_timer_read(vb200_ctx* ctx, double* ms_total, uint64_t* launches) {
    if (!ctx || !ms_total || !launches) return ctx ? vb200::fail(ctx, VB200_ERR_INVALID, "vb200_kernel_timer_read: null output") : VB200_ERR_INVALID;
    VB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double ms = 0.0;
    for (auto& e : ctx->ktimer_events) { float t = 0.f; VB200_CUDA(ctx, cudaEventElapsedTime(&t, e.first, e.second)); ms += t; }
    *ms_total = ms; *launches = ctx->ktimer_events.size();
    ktimer_drop(ctx);
    return VB200_OK;
}

extern "C" int vb200_host_register(vb200_ctx* ctx, void* ptr, size_t bytes) {
    if (!ctx || !ptr || bytes == 0) return fail(ctx, VB200_ERR_INVALID, "NULL/empty buffer");
    VB200_CUDA(ctx, cudaSetDevice(ctx->device));
    for (auto& r : ctx->registered) if (r.first == static_cast<char*>(ptr)) return fail(ctx, VB200_ERR_INVALID, "buffer %p is already registered", ptr);
    VB200_CUDA(ctx, cudaHostRegister(ptr, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
    ctx->registered.emplace_back(static_cast<char*>(ptr), bytes);
    return VB200_OK;
}
extern "C" int vb200_host_unregister(vb200_ctx* ctx, void* ptr) {
    if (!ctx || !ptr) return fail(ctx, VB200_ERR_INVALID, "NULL argument");
    VB200_CUDA(ctx, cudaSetDevice(ctx->device));
    for (size_t i = 0; i < ctx->registered.size(); ++i) if (ctx->registered[i].first == static_cast<char*>(ptr)) {
        VB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        VB200_CUDA(ctx, cudaHostUnregister(ptr));
        ctx->registered.erase(ctx->registered.begin() + long(i));
        return VB200_OK;
    }
    return fail(ctx, VB200_ERR_INVALID, "buffer %p is not registered", ptr);
}
namespace {
// device-visible alias of a host range if it lies inside a registered buffer, else nullptr
float* mapped_alias(vb200_ctx* ctx, float* p, size_t bytes) {
    const char* b = reinterpret_cast<const char*>(p);
    for (auto& r : ctx->registered) if (b >= r.first && b + bytes <= r.first + r.second) {
        void* d = nullptr;
        if (cudaHostGetDevicePointer(&d, p, 0) == cudaSuccess) return static_cast<float*>(d);
        cudaGetLastError();
    }
    return nullptr;
}
}

namespace {
__global__ void __launch_bounds__(256) fma_chain_kernel(float* out, int iters, float a, float b) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = float(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], a, b);
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    if (s == 12345.678f) out[0] = s;      // never true: keeps the chain alive
}
}

extern "C" int vb200_measure_fp32_peak(vb200_ctx* ctx, int reps, double* tflops) {
    if (!ctx || !tflops) return fail(ctx, VB200_ERR_INVALID, "NULL argument");
    VB200_CUDA(ctx, cudaSetDevice(ctx->device));
    void* scratch = nullptr; int rc = reserve(ctx, 3, 256, &scratch); if (rc) return rc;
    cudaEvent_t e0, e1; VB200_CUDA(ctx, cudaEventCreate(&e0)); VB200_CUDA(ctx, cudaEventCreate(&e1));
    const int iters = 8192, grid = ctx->sm_count * 8;
    float best = 1e30f;
    for (int r = 0; r < (reps < 1 ? 1 : reps) + 1; ++r) {
        cudaEventRecord(e0, ctx->stream);
        fma_chain_kernel<<<grid, 256, 0, ctx->stream>>>(static_cast<float*>(scratch), iters, 1.0001f, 0.5f);
        cudaEventRecord(e1, ctx->stream);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { cudaEventDestroy(e0); cudaEventDestroy(e1); return fail(ctx, VB200_ERR_CUDA, "FMA peak kernel failed: %s", cudaGetErrorString(e)); }
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;      // first launch is warm-up
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *tflops = double(grid) * 256.0 * double(iters) * 8.0 * 2.0 / (double(best) * 1e-3) / 1e12;
    return VB200_OK;
}

extern "C" void vb200_philox4x32_10(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]) {
    const viltrum::b200::u32x4 r = viltrum::b200::philox4x32<10>(viltrum::b200::u32x4{counter[0], counter[1], counter[2], counter[3]}, key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

extern "C" void vb200_xoshiro128pp(uint32_t state[4], uint64_t n, uint32_t* out) {
    viltrum::b200::Xoshiro128pp x{state[0], state[1], state[2], state[3]};
    for (uint64_t i = 0; i < n; ++i) out[i] = x.next();
    state[0] = x.s0; state[1] = x.s1; state[2] = x.s2; state[3] = x.s3;
}
extern "C" int vb200_threefry4x32(int rounds, const uint32_t counter[4], const uint32_t key[4], uint32_t out[4]) {
    using namespace viltrum::b200;
    const ThreefryKeys t = threefry_key_schedule(key[0], key[1], key[2], key[3]);
    const u32x4 c{counter[0], counter[1], counter[2], counter[3]};
    u32x4 r;
    if (rounds == 12) r = threefry4x32<12>(c, t); else if (rounds == 13) r = threefry4x32<13>(c, t); else if (rounds == 20) r = threefry4x32<20>(c, t);
    else return VB200_ERR_INVALID;
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
    return VB200_OK;
}

// ---- built-in integrands ----------------------------------------------------------------------------------
struct BuiltinEntry { const char* name; const vb200_integrand* desc; };
extern "C" const BuiltinEntry* builtin_table_fast(int* count);
extern "C" const BuiltinEntry* builtin_table_exact(int* count);

extern "C" const vb200_integrand* vb200_builtin_integrand(const char* name, int exact) {
    if (!name) return nullptr;
    int n = 0; const BuiltinEntry* t = exact ? builtin_table_exact(&n) : builtin_table_fast(&n);
    for (int i = 0; i < n; ++i) if (!std::strcmp(t[i].name, name)) return t[i].desc;
    return nullptr;
}
extern "C" const BuiltinEntry* builtin_table64_fast(int* count);
extern "C" const BuiltinEntry* builtin_table64_exact(int* count);
extern "C" const vb200_integrand* vb200_builtin_integrand_f64(const char* name, int exact) {
    if (!name) return nullptr;
    int n = 0; const BuiltinEntry* t = exact ? builtin_table64_exact(&n) : builtin_table64_fast(&n);
    for (int i = 0; i < n; ++i) if (!std::strcmp(t[i].name, name)) return t[i].desc;
    return nullptr;
}
extern "C" int vb200_builtin_count(void) { int n = 0; builtin_table_fast(&n); return n; }
extern "C" const char* vb200_builtin_name(int index) { int n = 0; const BuiltinEntry* t = builtin_table_fast(&n); return (index >= 0 && index < n) ? t[index].name : nullptr; }

// ---- per-bin Monte Carlo ----------------------------------------------------------------------------------
namespace {
// Range::volume(): float product over the dimensions in order, starting from 1 (reference src/range.h:21-25)
float range_volume(const vb200_domain& d, int n) { float v = 1.0f; for (int i = 0; i < n; ++i) v *= (d.rmax[i] - d.rmin[i]); return v; }

}

// Shared tail of the two sampling drivers (finite and infinite ranges).
//   DEVICE bins: one launch over the shard, '+=' (or '=') applied by the kernel in place; returns with work enqueued.
//   HOST bins  : the end-to-end path.  ONE launch writes its estimates straight into the context's pinned, device-mapped
//                staging buffer (zero-copy stores ride PCIe underneath the compute — no separate D2H pass) and raises a
//                host-mapped flag per chunk of consecutive tiles (vb200_chunk_signal); the host spins on the flags and applies
//                '+=' / '=' to chunk k while the kernel is still computing chunk k+1.  No events, no per-chunk launches: the
//                only serial tail is the last chunk's host pass.
template<class Launch>
static int run_sampler(vb200_ctx* ctx, const vb200_integrand* f, int kind, Launch& a, bool accumulate,
                       float* bins, int bins_mem, float* sum_f, float* sum_f2) {
    const uint64_t begin = a.bin_begin, end = a.bin_end, n = end - begin;
    if (bins_mem != VB200_HOST && bins_mem != VB200_DEVICE) return fail(ctx, VB200_ERR_INVALID, "bad memory-space flag %d", bins_mem);
    if (!bins) return fail(ctx, VB200_ERR_INVALID, "bins pointer is NULL");
    // the finite sampler resets its scheduler words itself (mc_per_bin.cuh); the walk kernels still want them cleared, and so does the first
    // call after a failed launch
    if (kind != VB200_K_MC_PER_BIN || ctx->counters_dirty) {
        VB200_CUDA(ctx, cudaMemsetAsync(ctx->d_counter, 0, 2 * sizeof(unsigned long long) + vb200_ctx::kMaxChunks * sizeof(uint32_t), ctx->stream));
        ctx->counters_dirty = false;
    }
    if (kind != VB200_K_MC_PER_BIN) ctx->counters_dirty = true;      // walk kernels leave their tickets behind
    a.tile_counter = ctx->d_counter;
    std::memset(&a.signal, 0, sizeof(a.signal));
    if (bins_mem == VB200_DEVICE) {
        a.accumulate = accumulate ? 1 : 0; a.out = bins; a.sum_f = sum_f; a.sum_f2 = sum_f2;
        return call_thunk(ctx, f, kind, &a);
    }
    if (!sum_f && !sum_f2) {
        // registered (pinned + mapped) caller bins: the kernel applies '+=' / '=' in place over PCIe; nothing left for the host to do
        if (float* alias = mapped_alias(ctx, bins + begin, n * sizeof(float))) {
            a.accumulate = accumulate ? 1 : 0; a.out = alias - begin; a.sum_f = nullptr; a.sum_f2 = nullptr;
            int rc0 = call_thunk(ctx, f, kind, &a); if (rc0) return rc0;
            VB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            return VB200_OK;
        }
    }
    float* h = nullptr;
    int rc = reserve_pinned(ctx, n * sizeof(float) * 3, reinterpret_cast<void**>(&h)); if (rc) return rc;
    a.accumulate = 0;
    a.out = h - begin;                                   // kernels index from the base of the full grid
    a.sum_f = sum_f ? h + n : nullptr; a.sum_f2 = sum_f2 ? h + 2 * n : nullptr;
    // chunking: tiles of G = 32/lanes_per_bin bins; chunks of 2^shift tiles, about 16 Ki bins each (the host pass over one chunk is
    // ~15 us, the exposed tail), at most kMaxChunks of them.  VB200_E2E_CHUNK_BINS overrides the target (tuning knob).
    const uint64_t G = 32u / a.lanes_per_bin, ntiles = (n + G - 1) / G;
    uint64_t target_bins = 16u * 1024u;                 // measured on the pool's 16-vCPU host (profiles/results_r1.md, round 1d): one host thread adds ~1 bin/ns, so a 64 Ki chunk is a 60 us tail
    if (const char* env = std::getenv("VB200_E2E_CHUNK_BINS")) { const long long v = std::atoll(env); if (v > 0) target_bins = uint64_t(v); }
    uint32_t shift = 0;
    while ((G << (shift + 1)) <= target_bins) ++shift;
    while (((ntiles + (1ull << shift) - 1) >> shift) > uint64_t(vb200_ctx::kMaxChunks)) ++shift;
    const uint64_t chunks = (ntiles + (1ull << shift) - 1) >> shift;
    a.signal.enabled = 1; a.signal.chunk_shift = shift; a.signal.done = ctx->d_done; a.signal.flag = ctx->h_flags;
    if (++ctx->epoch == 0) ++ctx->epoch;                 // never 0: the flags start out zeroed
    const uint32_t epoch = a.signal.epoch = ctx->epoch;
    rc = call_thunk(ctx, f, kind, &a); if (rc) return rc;
    HostPass job;
    job.flags = ctx->h_flags; job.epoch = epoch; job.chunks = chunks; job.bins_per_chunk = G << shift; job.n = n;
    job.dst = bins + begin; job.src = h; job.accumulate = accumulate;
    // helpers: VB200_HOST_THREADS (total threads incl. the caller's), default 8 or the machine's hardware threads if fewer
    // (measured on the pool's 16-vCPU host, C2: 1 thread 0.67 ms per call, 2: 0.51, 4: 0.42, 8: 0.405; kernel alone 0.343)
    int threads = 8;
    int hw = int(std::thread::hardware_concurrency());
    // one process per GPU: share the host's cores between the ranks of this node (torchrun exports LOCAL_WORLD_SIZE) — spinning
    // helpers must never outnumber the cores
    if (const char* lws = std::getenv("LOCAL_WORLD_SIZE")) { const int n = std::atoi(lws); if (n > 1 && hw > 0) hw = hw / n > 1 ? hw / n : 1; }
    if (hw > 0 && threads > hw) threads = hw;
    if (const char* env = std::getenv("VB200_HOST_THREADS")) threads = std::atoi(env);
    if (uint64_t(threads) > chunks) threads = int(chunks);
    if (threads < 1) threads = 1;
    // VB200_E2E_TRACE=1: one line of timestamps per call on stderr (us since the launch returned) — where the end-to-end time goes
    static const bool trace = std::getenv("VB200_E2E_TRACE") != nullptr;
    const auto t_launch = std::chrono::steady_clock::now();
    auto us = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::micro>(t - t_launch).count(); };
    if (threads > 1) { ctx->pool.start(threads - 1); ctx->pool.publish(&job); }
    const auto t_pub = std::chrono::steady_clock::now();
    cudaError_t stream_error = cudaSuccess;
    job.run(true, ctx->stream, &stream_error);
    const auto t_run = std::chrono::steady_clock::now();
    uint64_t spins = 0;
    while (job.finished.load(std::memory_order_acquire) != chunks && !job.abort.load()) {
        if ((++spins & 0x3fffu) == 0) {
            const cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q != cudaSuccess && q != cudaErrorNotReady) { stream_error = q; job.abort.store(1); }
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    const auto t_fin = std::chrono::steady_clock::now();
    if (threads > 1) ctx->pool.retire();
    const auto t_ret = std::chrono::steady_clock::now();
    if (job.abort.load()) { cudaStreamSynchronize(ctx->stream); return fail(ctx, VB200_ERR_CUDA, "sampling kernel failed or ended without completing its chunks: %s",
            cudaGetErrorString(stream_error)); }
    VB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (trace) {
        const auto t_sync = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[vb200 e2e] chunks %llu x %llu bins, %d threads: published %.1f, own pass done %.1f, all chunks done %.1f, pool retired %.1f, stream synced %.1f us\n",
                     (unsigned long long)chunks, (unsigned long long)(G << shift), threads, us(t_pub), us(t_run), us(t_fin), us(t_ret), us(t_sync));
    }
    if (sum_f)  std::memcpy(sum_f, h + n, n * sizeof(float));
    if (sum_f2) std::memcpy(sum_f2, h + 2 * n, n * sizeof(float));
    return VB200_OK;
}

extern "C" int vb200_mc_per_bin(vb200_ctx* ctx, const vb200_integrand* f, const vb200_mc_params* p,
                                float* bins, int bins_mem, float* sum_f, float* sum_f2) {
    if (!ctx || !f || !p) return fail(ctx, VB200_ERR_INVALID, "NULL argument");
    VB200_CUDA(ctx, cudaSetDevice(ctx->device));
    if (f->dim <= 0) return fail(ctx, VB200_ERR_INVALID, "vb200_mc_per_bin needs a finite-dimensional integrand (use vb200_mc_per_bin_inf)");
    int rc = check_domain(ctx, p->domain, f->dim); if (rc) return rc;
    if (p->spp == 0 || p->spp > 0xffffffffull) return fail(ctx, VB200_ERR_INVALID, "spp=%llu invalid", (unsigned long long)p->spp);
    if (p->flavor != VB200_MC_PER_BIN && p->flavor != VB200_PER_BIN_MC) return fail(ctx, VB200_ERR_INVALID, "unknown flavor %d", p->flavor);
    const uint64_t total = nbins_of(p->domain);
    uint64_t begin, end; rc = resolve_shard(ctx, p->shard, total, &begin, &end); if (rc) return rc;
    if (begin == end) return VB200_OK;

    vb200_mc_launch a; std::memset(&a, 0, sizeof(a));
    a.domain = finish_domain(p->domain); a.bin_begin = begin; a.bin_end = end; a.nbins_total = total;
    a.spp = uint32_t(p->spp); a.lanes_per_bin = pick_lanes_per_bin(ctx, p->spp, total, 8);   // from the WHOLE grid: the summation order, hence the bits, must not depend on the shard
    if (p->options & ~(VB200_MC_RNG_PHILOX | VB200_MC_LATTICE24)) return fail(ctx, VB200_ERR_INVALID, "unknown option bits 0x%x", unsigned(p->options));
    a.rng = (p->options & VB200_MC_RNG_PHILOX) ? VB200_RNG_PHILOX : VB200_RNG_XOSHIRO;
    a.narrow_binned = (p->options & VB200_MC_LATTICE24) ? 0 : 1;      // 16-bit draws inside a bin only where the grid itself supplies >= 8 bits per binned dimension
    for (int i = 0; i < p->domain.dimbins; ++i) if (p->domain.res[i] < 256) a.narrow_binned = 0;
    a.key0 = uint32_t(p->seed); a.key1 = uint32_t(p->seed >> 32);
    a.flavor = p->flavor;
    a.factor = double(range_volume(p->domain, p->domain.dim)) / double(p->spp);     // monte-carlo-per-bin-parallel.h:45
    // '+=' for MonteCarloPerBinParallel, '=' for IntegratorPerBinParallel
    return run_sampler(ctx, f, VB200_K_MC_PER_BIN, a, p->flavor == VB200_MC_PER_BIN, bins, bins_mem, sum_f, sum_f2);
}

extern "C" int vb200_mc_per_bin_replay(vb200_ctx* ctx, const vb200_integrand* f, const vb200_mc_params* p,
                                       const float* samples, int samples_mem, float* bins, int bins_mem) {
    if (!ctx || !f || !p || !samples) return fail(ctx, VB200_ERR_INVALID, "NULL argument");
    VB200_CUDA(ctx, cudaSetDevice(ctx->device));
    if (f->dim <= 0) return fail(ctx, VB200_ERR_INVALID, "replay of finite samples needs a finite-dimensional integrand");
    int rc = check_domain(ctx, p->domain, f->dim); if (rc) return rc;
    if (p->spp == 0 || p->spp > 0xffffffffull) return fail(ctx, VB200_ERR_INVALID, "spp invalid");
    const uint64_t total = nbins_of(p->domain);
    uint64_t begin, end; rc = resolve_shard(ctx, p->shard, total, &begin, &end); if (rc) return rc;
    if (begin == end) return VB200_OK;
    vb200_replay_launch a; std::memset(&a, 0, sizeof(a));
    a.domain = finish_domain(p->domain); a.bin_begin = begin; a.bin_end = end; a.nbins_total = total; a.spp = uint32_t(p->spp); a.flavor = p->flavor;
    a.factor = double(range_volume(p->domain, p->domain.dim)) / double(p->spp);
    const size_t sbytes = size_t(end - begin) * p->spp * size_t(f->dim) * sizeof(float);
    if (samples_mem == VB200_HOST) {
        void* d = nullptr; rc = reserve(ctx, 1, sbytes, &d); if (rc) return rc;
        VB200_CUDA(ctx, cudaMemcpyAsync(d, samples, sbytes, cudaMemcpyHostToDevice, ctx->stream));
        a.samples = static_cast<const float*>(d);
    } else a.samples = samples;
    // replay continues the reference's running '+=' from the bins' current contents: upload them
    BinStage st; rc = stage_bins_in(ctx, bins, bins_mem, begin, end, /*upload=*/true, &st); if (rc) return rc;
    a.out = st.dev_base;
    rc = call_thunk(ctx, f, VB200_K_MC_REPLAY, &a); if (rc) return rc;
    return stage_bins_out(ctx, st);
}

// ---- infinite-dimensional paths ---------------------------------------------------------------------------
extern "C" int vb200_mc_per_bin_inf(vb200_ctx* ctx, const vb200_integrand* f, const vb200_mc_params* p,
                                    float* bins, int bins_mem, float* sum_f, float* sum_f2) {
    if (!ctx || !f || !p) return fail(ctx, VB200_ERR_INVALID, "NULL argument");
    VB200_CUDA(ctx, cudaSetDevice(ctx->device));
    if (f->dim != -1) return fail(ctx, VB200_ERR_INVALID, "vb200_mc_per_bin_inf needs a sequence integrand");
    int rc = check_domain(ctx, p->domain, -1); if (rc) return rc;
    if (p->spp == 0 || p->spp > 0xffffffffull) return fail(ctx, VB200_ERR_INVALID, "spp invalid");
    if (p->flavor != VB200_MC_PER_BIN && p->flavor != VB200_PER_BIN_MC) return fail(ctx, VB200_ERR_INVALID, "unknown flavor %d", p->flavor);
    const uint64_t total = nbins_of(p->domain);
    uint64_t begin, end; rc = resolve_shard(ctx, p->shard, total, &begin, &end); if (rc) return rc;
    if (begin == end) return VB200_OK;
    vb200_walk_launch a; std::memset(&a, 0, sizeof(a));
    a.domain = finish_domain(p->domain); a.bin_begin = begin; a.bin_end = end; a.nbins_total = total;
    a.spp = uint32_t(p->spp); a.lanes_per_bin = pick_lanes_per_bin(ctx, p->spp, total, 4);   // from the WHOLE grid: the summation order, hence the bits, must not depend on the shard
    a.key0 = uint32_t(p->seed); a.key1 = uint32_t(p->seed >> 32);
    a.factor = double(range_volume(p->domain, p->domain.dim)) / double(p->spp);      // range-infinite.h:22-23; :77
    a.flavor = p->flavor;
    // '+=' for MonteCarloPerBinParallel, '=' for IntegratorPerBinParallel(MonteCarlo)
    return run_sampler(ctx, f, VB200_K_WALK, a, p->flavor == VB200_MC_PER_BIN, bins, bins_mem, sum_f, sum_f2);
}

extern "C" int vb200_mc_per_bin_inf_replay(vb200_ctx* ctx, const vb200_integrand* f, const vb200_mc_params* p,
                                           const uint64_t* offsets, const float* elems, int mem, float* bins, int bins_mem) {
    if (!ctx || !f || !p || !offsets || !elems) return fail(ctx, VB200_ERR_INVALID, "NULL argument");
    VB200_CUDA(ctx, cudaSetDevice(ctx->device));
    if (f->dim != -1) return fail(ctx, VB200_ERR_INVALID, "sequence replay needs a sequence integrand");
    int rc = check_domain(ctx, p->domain, -1); if (rc) return rc;
    const uint64_t total = nbins_of(p->domain);
    if (p->spp == 0 || p->spp > 0xffffffffull) return fail(ctx, VB200_ERR_INVALID, "spp invalid");
    if (p->flavor != VB200_MC_PER_BIN && p->flavor != VB200_PER_BIN_MC) return fail(ctx, VB200_ERR_INVALID, "unknown flavor %d", p->flavor);
    uint64_t begin, end; rc = resolve_shard(ctx, p->shard, total, &begin, &end); if (rc) return rc;
    if (begin == end) return VB200_OK;
    const uint64_t npaths = (end - begin) * p->spp;
    vb200_walk_replay_launch a; std::memset(&a, 0, sizeof(a));
    a.domain = finish_domain(p->domain); a.bin_begin = begin; a.bin_end = end; a.nbins_total = total; a.spp = uint32_t(p->spp); a.flavor = p->flavor;
    a.factor = double(range_volume(p->domain, p->domain.dim)) / double(p->spp);
    if (mem == VB200_HOST) {
        const uint64_t nel = offsets[npaths];
        void *d0 = nullptr, *d1 = nullptr;
        rc = reserve(ctx, 1, (npaths + 1) * sizeof(uint64_t), &d0); if (rc) return rc;
        rc = reserve(ctx, 2, (nel + 1) * sizeof(float), &d1); if (rc) return rc;
        VB200_CUDA(ctx, cudaMemcpyAsync(d0, offsets, (npaths + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        VB200_CUDA(ctx, cudaMemcpyAsync(d1, elems, nel * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        a.offsets = static_cast<const uint64_t*>(d0); a.elems = static_cast<const float*>(d1);
    } else { a.offsets = offsets; a.elems = elems; }
    VB200_CUDA(ctx, cudaMemsetAsync(ctx->d_flag, 0, sizeof(int32_t), ctx->stream));
    a.error_flag = ctx->d_flag;
    BinStage st; rc = stage_bins_in(ctx, bins, bins_mem, begin, end, true, &st); if (rc) return rc;
    a.out = st.dev_base;
    rc = call_thunk(ctx, f, VB200_K_WALK_REPLAY, &a); if (rc) return rc;
    rc = stage_bins_out(ctx, st); if (rc) return rc;
    int32_t flag = 0;
    VB200_CUDA(ctx, cudaMemcpyAsync(&flag, ctx->d_flag, sizeof(flag), cudaMemcpyDeviceToHost, ctx->stream));
    VB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (flag) return fail(ctx, VB200_ERR_INVALID, "sequence replay: the integrand read past the recorded length of a path");
    return VB200_OK;
}

// bins[i] = float(double(bins[i]) + double(add[i])): the reference's '+=' onto device-resident bins
__global__ void add_into_kernel(float* __restrict__ bins, const float* __restrict__ add, uint64_t n) {
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x)
        bins[i] = __double2float_rn(__dadd_rn(double(bins[i]), double(add[i])));
}
namespace vb200 {
int add_into(vb200_ctx* ctx, float* bins, const float* add, uint64_t n) {
    const unsigned grid = unsigned(std::min<uint64_t>((n + 255) / 256, uint64_t(ctx->sm_count) * 8));
    add_into_kernel<<<grid, 256, 0, ctx->stream>>>(bins, add, n);
    ctx->launches++;
    VB200_CUDA(ctx, cudaGetLastError());
    return VB200_OK;
}
}

// ---- global Monte Carlo (scatter) ---------------------------------------------------------------------------
extern "C" int vb200_monte_carlo(vb200_ctx* ctx, const vb200_integrand* f, const vb200_mc_params* p, float* bins, int bins_mem) {
    if (!ctx || !f || !p) return fail(ctx, VB200_ERR_INVALID, "NULL argument");
    VB200_CUDA(ctx, cudaSetDevice(ctx->device));
    if (f->dim == 0) return fail(ctx, VB200_ERR_INVALID, "integrand has dimension 0");
    int rc = check_domain(ctx, p->domain, f->dim > 0 ? f->dim : -1); if (rc) return rc;
    if (p->spp == 0) return fail(ctx, VB200_ERR_INVALID, "samples=0");
    const uint64_t total = nbins_of(p->domain);
    if (p->options & ~VB200_MC_ALLREDUCE) return fail(ctx, VB200_ERR_INVALID, "unknown option bits 0x%x", unsigned(p->options));
    const bool allreduce = (p->options & VB200_MC_ALLREDUCE) != 0;
    uint64_t sb, se;
    if (allreduce) {      // split-sample mode: this rank's share of the global sample counter; the partial grids are summed over the communicator
        if (!ctx->comm) return fail(ctx, VB200_ERR_INVALID, "VB200_MC_ALLREDUCE needs a communicator: call vb200_comm_init first");
        if (p->shard.begin != 0 || p->shard.end != 0) return fail(ctx, VB200_ERR_INVALID, "VB200_MC_ALLREDUCE derives the sample shard from the communicator: leave shard {0,0}");
        sb = uint64_t((unsigned __int128)p->spp * unsigned(ctx->comm_rank) / unsigned(ctx->comm_size));          // = host.py sample_shard_for_rank
        se = uint64_t((unsigned __int128)p->spp * unsigned(ctx->comm_rank + 1) / unsigned(ctx->comm_size));
    } else { rc = resolve_shard(ctx, p->shard, p->spp, &sb, &se); if (rc) return rc; }
    vb200_scatter_launch a; std::memset(&a, 0, sizeof(a));
    a.domain = finish_domain(p->domain); a.sample_begin = sb; a.sample_end = se; a.nbins_total = total;
    a.key0 = uint32_t(p->seed); a.key1 = uint32_t(p->seed >> 32);
    a.factor = double(total) * double(range_volume(p->domain, p->domain.dim)) / double(p->spp);     // monte-carlo.h:43-45
    float* dev = bins;
    if (bins_mem == VB200_HOST || allreduce) {      // a zeroed partial grid: the '+=' onto the caller's bins comes after the sum over ranks
        void* d = nullptr; rc = reserve(ctx, 0, total * sizeof(float), &d); if (rc) return rc;
        dev = static_cast<float*>(d);
        VB200_CUDA(ctx, cudaMemsetAsync(dev, 0, total * sizeof(float), ctx->stream));
    }
    a.out = dev;
    if (se > sb) { rc = call_thunk(ctx, f, f->dim > 0 ? VB200_K_MC_SCATTER : VB200_K_WALK_SCATTER, &a); if (rc) return rc; }
    if (allreduce) {
        rc = comm_allreduce_sum(ctx, dev, total); if (rc) return rc;
        if (bins_mem == VB200_DEVICE) { rc = vb200::add_into(ctx, bins, dev, total); if (rc) return rc; }
    }
    if (bins_mem == VB200_HOST) {
        float* h = nullptr; rc = reserve_pinned(ctx, total * sizeof(float), reinterpret_cast<void**>(&h)); if (rc) return rc;
        VB200_CUDA(ctx, cudaMemcpyAsync(h, dev, total * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        VB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (uint64_t i = 0; i < total; ++i) bins[i] = float(double(bins[i]) + double(h[i]));        // '+=' (monte-carlo.h:59)
    }
    return VB200_OK;
}
