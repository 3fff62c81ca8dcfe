// Region tables and region->bin integration (kernels K8-K10 of §2.2).
// Compiled with --fmad=false; the rule arithmetic additionally uses explicit round-to-nearest intrinsics
// (include/viltrum_b200/device/rules.cuh), so every float/double rounding of the reference happens here too and the
// per-bin results are bit-identical to RegionsIntegratorSequential on the same region table.
#include "regions.h"
#include <viltrum_b200/device/rules.cuh>
#include <vector>
#include <cstring>
#include <cstdlib>
#include <new>

using namespace vb200;
namespace R = viltrum::b200::device::rules;

namespace vb200 {

int rule_samples(int rule, int* SH, int* SL) {
    if (VB200_RULE_IS_STEPS(rule)) {       // Steps<Q,N>: samples = (Q::samples - 1)*N + 1 (rules.h:326)
        const int q = (rule >> 16) & 0xff, n = rule & 0xffff;
        if ((q != 2 && q != 3 && q != 5) || n < 1) return -1;
        *SH = (q - 1) * n + 1; *SL = 0; return 0;
    }
    switch (rule) {
        case VB200_RULE_TRAPEZOIDAL: *SH = 2; *SL = 0; return 0;
        case VB200_RULE_SIMPSON: *SH = 3; *SL = 0; return 0;
        case VB200_RULE_BOOLE: *SH = 5; *SL = 0; return 0;
        case VB200_RULE_SIMPSON_TRAPEZOIDAL: *SH = 3; *SL = 2; return 0;
        case VB200_RULE_BOOLE_SIMPSON: *SH = 5; *SL = 3; return 0;
    }
    return -1;
}

int regions_alloc(vb200_ctx* ctx, int dim, int rule, uint64_t capacity, vb200_regions** out, bool f64) {
    int SH, SL;
    if (rule_samples(rule, &SH, &SL)) return fail(ctx, VB200_ERR_INVALID, "unknown rule %d", rule);
    if (dim < 1 || dim > VB200_MAX_DIM) return fail(ctx, VB200_ERR_INVALID, "region dimension %d outside 1..%d", dim, VB200_MAX_DIM);
    uint64_t sd = 1; for (int i = 0; i < dim; ++i) { sd *= uint64_t(SH); if (sd > (1ull << 40)) break; }
    if (VB200_RULE_IS_STEPS(rule) ? (sd > (1ull << 22) || capacity != 1) : sd > 15625) return fail(ctx, VB200_ERR_UNSUPPORTED,
            "%llu samples per region: beyond the reference's own limit (VILTRUM_MAX_DIMENSIONS_REGION, region.h:16-18)", (unsigned long long)sd);
    vb200_regions* r = new (std::nothrow) vb200_regions;
    if (!r) return fail(ctx, VB200_ERR_NOMEM, "out of host memory");
    r->ctx = ctx; r->dim = dim; r->rule = rule; r->SH = SH; r->SL = SL; r->sd = int(sd); r->capacity = capacity; r->count = 0; r->f64 = f64;
    ctx->live_regions.push_back(r);
    cudaError_t e;
    if (f64 ? ((e = dmalloc(ctx, &r->rmin64, capacity * dim * sizeof(double))) != cudaSuccess || (e = dmalloc(ctx, &r->rmax64, capacity * dim * sizeof(double))) != cudaSuccess ||
               (e = dmalloc(ctx, &r->data64, capacity * sd * sizeof(double))) != cudaSuccess || (e = dmalloc(ctx, &r->err64, capacity * sizeof(double))) != cudaSuccess ||
               (e = dmalloc(ctx, &r->errdim, capacity * sizeof(uint32_t))) != cudaSuccess)
            : ((e = dmalloc(ctx, &r->rmin, capacity * dim * sizeof(float))) != cudaSuccess || (e = dmalloc(ctx, &r->rmax, capacity * dim * sizeof(float))) != cudaSuccess ||
        (e = dmalloc(ctx, &r->data, capacity * sd * sizeof(float))) != cudaSuccess || (e = dmalloc(ctx, &r->err, capacity * sizeof(float))) != cudaSuccess ||
        (e = dmalloc(ctx, &r->errdim, capacity * sizeof(uint32_t))) != cudaSuccess)) {
        cudaGetLastError(); vb200_regions_free(r);
        return fail(ctx, VB200_ERR_NOMEM, "region table of %llu regions x %llu samples does not fit: %s", (unsigned long long)capacity, (unsigned long long)sd, cudaGetErrorString(e));
    }
    *out = r;
    return VB200_OK;
}

} // namespace vb200

namespace vb200 {
// called by vb200_destroy: the context is going away — release the device memory of every outstanding table and cut the link
void orphan_regions(vb200_ctx* ctx) {
    for (vb200_regions* r : ctx->live_regions) {
        cudaFree(r->rmin); cudaFree(r->rmax); cudaFree(r->data); cudaFree(r->err); cudaFree(r->errdim);
        cudaFree(r->rmin64); cudaFree(r->rmax64); cudaFree(r->data64); cudaFree(r->err64);
        r->rmin = r->rmax = r->data = r->err = nullptr; r->errdim = nullptr; r->rmin64 = r->rmax64 = r->data64 = r->err64 = nullptr;
        r->ctx = nullptr; r->count = 0;
    }
    ctx->live_regions.clear();
}
}

extern "C" void vb200_regions_free(vb200_regions* r) {
    if (!r) return;
    vb200_ctx* ctx = r->ctx;       // stream-ordered frees: work already enqueued on the context's stream still sees the table
    if (!ctx) { delete r; return; }      // the context was destroyed first: vb200_destroy already released the device memory
    for (size_t i = 0; i < ctx->live_regions.size(); ++i) if (ctx->live_regions[i] == r) { ctx->live_regions[i] = ctx->live_regions.back(); ctx->live_regions.pop_back(); break; }
    dfree(ctx, r->rmin); dfree(ctx, r->rmax); dfree(ctx, r->data); dfree(ctx, r->err); dfree(ctx, r->errdim);
    dfree(ctx, r->rmin64); dfree(ctx, r->rmax64); dfree(ctx, r->data64); dfree(ctx, r->err64);
    delete r;
}
extern "C" uint64_t vb200_regions_count(const vb200_regions* r) { return r ? r->count : 0; }
extern "C" int vb200_regions_dim(const vb200_regions* r) { return r ? r->dim : 0; }
extern "C" int vb200_regions_samples(const vb200_regions* r) { return r ? r->sd : 0; }

template<class T>
static int regions_upload_t(vb200_ctx* ctx, int dim, int rule, uint64_t count, const T* rmin, const T* rmax, const T* err, const uint32_t* errdim, const T* data, vb200_regions** out) {
    if (!ctx || !out || !rmin || !rmax || !data || count == 0) return fail(ctx, VB200_ERR_INVALID, "NULL/empty argument");
    VB200_CUDA(ctx, cudaSetDevice(ctx->device));
    vb200_regions* r = nullptr;
    int rc = regions_alloc(ctx, dim, rule, count, &r, sizeof(T) == 8); if (rc) return rc;
    const uint64_t sd = uint64_t(r->sd);
    std::vector<T> t(count * (sd > uint64_t(dim) ? sd : uint64_t(dim)));
    auto up = [&] (T* dst, const T* src, uint64_t width) -> cudaError_t {       // AoS [count][width] -> SoA [width][count]
        for (uint64_t i = 0; i < count; ++i) for (uint64_t k = 0; k < width; ++k) t[k * count + i] = src[i * width + k];
        return cudaMemcpyAsync(dst, t.data(), count * width * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
    };
    cudaError_t e;
    if ((e = up(RegCols<T>::rmin(r), rmin, dim)) != cudaSuccess || (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess ||
        (e = up(RegCols<T>::rmax(r), rmax, dim)) != cudaSuccess || (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess ||
        (e = up(RegCols<T>::data(r), data, sd)) != cudaSuccess || (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) {
        vb200_regions_free(r); return fail(ctx, VB200_ERR_CUDA, "region upload failed: %s", cudaGetErrorString(e));
    }
    if ((e = err ? cudaMemcpyAsync(RegCols<T>::err(r), err, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream)
                 : cudaMemsetAsync(RegCols<T>::err(r), 0, count * sizeof(T), ctx->stream)) != cudaSuccess ||
        (e = errdim ? cudaMemcpyAsync(r->errdim, errdim, count * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream)
                    : cudaMemsetAsync(r->errdim, 0, count * sizeof(uint32_t), ctx->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) {
        vb200_regions_free(r); return fail(ctx, VB200_ERR_CUDA, "region upload failed: %s", cudaGetErrorString(e));
    }
    r->count = count;
    *out = r;
    return VB200_OK;
}
extern "C" int vb200_regions_upload(vb200_ctx* ctx, int dim, int rule, uint64_t count,
                                    const float* rmin, const float* rmax, const float* err, const uint32_t* errdim, const float* data, vb200_regions** out) {
    return regions_upload_t<float>(ctx, dim, rule, count, rmin, rmax, err, errdim, data, out);
}
extern "C" int vb200_regions_upload_f64(vb200_ctx* ctx, int dim, int rule, uint64_t count,
                                        const double* rmin, const double* rmax, const double* err, const uint32_t* errdim, const double* data, vb200_regions** out) {
    return regions_upload_t<double>(ctx, dim, rule, count, rmin, rmax, err, errdim, data, out);
}

template<class T>
static int regions_download_t(vb200_ctx* ctx, const vb200_regions* r, T* rmin, T* rmax, T* err, uint32_t* errdim, T* data) {
    if (!ctx || !r) return fail(ctx, VB200_ERR_INVALID, "NULL argument");
    if (r->f64 != (sizeof(T) == 8)) return fail(ctx, VB200_ERR_INVALID, "region table holds %s values", r->f64 ? "double" : "float");
    VB200_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t n = r->count, cap = r->capacity, sd = uint64_t(r->sd), dim = uint64_t(r->dim);
    std::vector<T> t;
    auto down = [&] (T* dst, const T* src, uint64_t width) -> int {              // SoA [width][cap] -> AoS [n][width]
        t.resize(width * cap);
        VB200_CUDA(ctx, cudaMemcpyAsync(t.data(), src, width * cap * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
        VB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (uint64_t i = 0; i < n; ++i) for (uint64_t k = 0; k < width; ++k) dst[i * width + k] = t[k * cap + i];
        return VB200_OK;
    };
    int rc;
    if (rmin && (rc = down(rmin, RegCols<T>::rmin(r), dim))) return rc;
    if (rmax && (rc = down(rmax, RegCols<T>::rmax(r), dim))) return rc;
    if (data && (rc = down(data, RegCols<T>::data(r), sd))) return rc;
    if (err) VB200_CUDA(ctx, cudaMemcpyAsync(err, RegCols<T>::err(r), n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    if (errdim) VB200_CUDA(ctx, cudaMemcpyAsync(errdim, r->errdim, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    VB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VB200_OK;
}
extern "C" int vb200_regions_download(vb200_ctx* ctx, const vb200_regions* r, float* rmin, float* rmax, float* err, uint32_t* errdim, float* data) {
    return regions_download_t<float>(ctx, r, rmin, rmax, err, errdim, data);
}
extern "C" int vb200_regions_download_f64(vb200_ctx* ctx, const vb200_regions* r, double* rmin, double* rmax, double* err, uint32_t* errdim, double* data) {
    return regions_download_t<double>(ctx, r, rmin, rmax, err, errdim, data);
}

// ---- regions_generator_single: one region over the whole range (regions-generator-single.h:12-20) --------------------
// sample points of region.h:40-46 + fill.h:45-72: p = double(i)/double(S-1); x = Float(p*(max-min) + min), (max-min) a difference in Float
template<class T>
__global__ void region_grid_points_kernel(int S, int dim, uint64_t n, const T* rmin, const T* rmax, uint64_t cap, uint64_t region, T* points) {
    const uint64_t k = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint64_t t = k;
    for (int d = 0; d < dim; ++d) {
        const double p = R::dd(double(t % uint64_t(S)), double(S - 1)); t /= uint64_t(S);
        const T lo = rmin[uint64_t(d) * cap + region], hi = rmax[uint64_t(d) * cap + region];
        points[uint64_t(d) * n + k] = R::from_double<T>(R::da(R::dm(p, double(R::sub(hi, lo))), double(lo)));
    }
}

template<class T>
static int generate_single_t(vb200_ctx* ctx, const vb200_integrand* f, int dim, const T* rmin, const T* rmax, int rule, vb200_regions** out) {
    VB200_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool f64 = sizeof(T) == 8;
    if (f->dim <= 0 || dim != f->dim) return fail(ctx, VB200_ERR_INVALID, "range has %d dimensions, integrand takes %d", dim, f->dim);
    if (bool(f->flags & VB200_INTEGRAND_F64) != f64) return fail(ctx, VB200_ERR_INVALID, "integrand '%s' computes in %s, the range is %s", f->name ? f->name : "?",
            (f->flags & VB200_INTEGRAND_F64) ? "double" : "float", f64 ? "double" : "float");
    vb200_regions* r = nullptr;
    int rc = regions_alloc(ctx, f->dim, rule, 1, &r, f64); if (rc) return rc;
    auto bail = [&] (int code) { vb200_regions_free(r); return code; };
    if (cudaMemcpyAsync(RegCols<T>::rmin(r), rmin, f->dim * sizeof(T), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
        cudaMemcpyAsync(RegCols<T>::rmax(r), rmax, f->dim * sizeof(T), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
        return bail(fail(ctx, VB200_ERR_CUDA, "range upload failed"));
    const uint64_t n = uint64_t(r->sd);
    void* pts = nullptr; rc = reserve(ctx, 3, n * f->dim * sizeof(T), &pts); if (rc) return bail(rc);
    region_grid_points_kernel<T><<<unsigned((n + 255) / 256), 256, 0, ctx->stream>>>(r->SH, f->dim, n, RegCols<T>::rmin(r), RegCols<T>::rmax(r), 1, 0, static_cast<T*>(pts));
    ctx->launches++;
    vb200_eval_launch ev; std::memset(&ev, 0, sizeof(ev));
    ev.n = n; ev.dim = f->dim; ev.f64 = f64 ? 1 : 0; ev.points = pts; ev.values = RegCols<T>::data(r);       // capacity 1: data[k*1+0]
    rc = call_thunk(ctx, f, VB200_K_EVAL_POINTS, &ev); if (rc) return bail(rc);
    if (cudaMemsetAsync(RegCols<T>::err(r), 0, sizeof(T), ctx->stream) != cudaSuccess || cudaMemsetAsync(r->errdim, 0, sizeof(uint32_t), ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) return bail(fail(ctx, VB200_ERR_CUDA, "single-region generation failed: %s", cudaGetErrorString(cudaGetLastError())));
    r->count = 1;
    *out = r;
    return VB200_OK;
}
extern "C" int vb200_regions_generate_single(vb200_ctx* ctx, const vb200_integrand* f, const vb200_domain* domain, int rule, vb200_regions** out) {
    if (!ctx || !f || !domain || !out) return fail(ctx, VB200_ERR_INVALID, "NULL argument");
    return generate_single_t<float>(ctx, f, domain->dim, domain->rmin, domain->rmax, rule, out);
}
extern "C" int vb200_regions_generate_single_f64(vb200_ctx* ctx, const vb200_integrand* f, const vb200_domain_f64* domain, int rule, vb200_regions** out) {
    if (!ctx || !f || !domain || !out) return fail(ctx, VB200_ERR_INVALID, "NULL argument");
    return generate_single_t<double>(ctx, f, domain->dim, domain->rmin, domain->rmax, rule, out);
}

// =====================================================================================================================
// Bin walk
// =====================================================================================================================
namespace {

// one fold level of Region::sub_last over a NON-binned dimension: the bin box spans the region's whole extent there, so the
// normalised limits are exactly 0 and 1 (pos_in_range(min)=0, pos_in_range(max)=1) and the result is the same for every
// bin — computed once per region instead of once per (bin, region) pair.  in: [S^m][cap] -> out: [S^(m-1)][cap].
template<int S, class T>
__global__ void fold_last_dim_kernel(uint64_t nregions, uint64_t cap, int lower /* S^(m-1) */, const T* __restrict__ in, T* __restrict__ out) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (i >= nregions) return;
    T line[S];
#pragma unroll
    for (int j = 0; j < S; ++j) line[j] = in[(uint64_t(k) + uint64_t(j) * uint64_t(lower)) * cap + i];
    out[uint64_t(k) * cap + i] = R::subrange<S, T>(T(0), T(1), line);
}

// the same fold level for Region::pdf_integral_subrange -> pdf_sub (region.h:277-302; Simpson only): the line integral of the
// shifted absolute parabola instead of the signed one
template<class T>
__global__ void pdf_fold_last_dim_kernel(uint64_t nregions, uint64_t cap, int lower, const T* __restrict__ in, T* __restrict__ out) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (i >= nregions) return;
    T line[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) line[j] = in[(uint64_t(k) + uint64_t(j) * uint64_t(lower)) * cap + i];
    out[uint64_t(k) * cap + i] = R::simpson_pdf_integral_subrange<T>(T(0), T(1), line);
}

// Range::volume (range.h:21-25) and pixels_in_region (region.h:454-463) per region
template<class T>
__global__ void region_boxes_kernel(uint64_t nregions, uint64_t cap, int dim, int db, DomT<T> dom,
                                    const T* __restrict__ rmin, const T* __restrict__ rmax,
                                    T* __restrict__ volume, uint32_t* __restrict__ pstart, uint32_t* __restrict__ pend) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nregions) return;
    T v = T(1);
    for (int d = 0; d < dim; ++d) v = R::mul(v, R::sub(rmax[uint64_t(d) * cap + i], rmin[uint64_t(d) * cap + i]));
    volume[i] = v;
    for (int d = 0; d < db; ++d) {
        const T res = T(dom.res[d]);
        const T ext = R::sub(dom.rmax[d], dom.rmin[d]);
        const T fs_ = R::quo(R::mul(res, R::sub(rmin[uint64_t(d) * cap + i], dom.rmin[d])), ext);
        const T fe_ = R::add(T(0.99f), R::quo(R::mul(res, R::sub(rmax[uint64_t(d) * cap + i], dom.rmin[d])), ext));     // 0.99f is a float literal upstream
        // size_t(Float): negative / NaN inputs are undefined upstream; clamp them to 0 here
        uint64_t s = fs_ > T(0) ? uint64_t(fs_) : 0ull;
        uint64_t e = fe_ > T(0) ? uint64_t(fe_) : 0ull;
        if (e > dom.res[d]) e = dom.res[d];
        if (e < s + 1) e = s + 1;
        pstart[uint64_t(d) * cap + i] = uint32_t(s < 0xffffffffull ? s : 0xffffffffull);
        pend[uint64_t(d) * cap + i] = uint32_t(e < 0xffffffffull ? e : 0xffffffffull);
    }
}

struct TileGeom { uint32_t tile[3], tiles[3], res[3]; int db; };

__device__ __forceinline__ void tile_origin(const TileGeom& g, uint64_t t, uint32_t (&o)[3]) {
    for (int d = 0; d < 3; ++d) { o[d] = uint32_t(t % g.tiles[d]) * g.tile[d]; t /= g.tiles[d]; }
}
__device__ __forceinline__ bool tile_in_shard(const TileGeom& g, const uint32_t (&o)[3], uint64_t begin, uint64_t end) {
    uint64_t lo = 0, hi = 0, prod = 1;
    for (int d = 0; d < g.db; ++d) {
        const uint32_t last = min(o[d] + g.tile[d], g.res[d]) - 1u;
        lo += uint64_t(o[d]) * prod; hi += uint64_t(last) * prod; prod *= g.res[d];
    }
    return hi >= begin && lo < end;
}
__device__ __forceinline__ bool overlaps(const TileGeom& g, const uint32_t (&o)[3], const uint32_t* pstart, const uint32_t* pend, uint64_t cap, uint64_t r) {
    bool ok = true;
    for (int d = 0; d < g.db; ++d) {
        const uint32_t s = pstart[uint64_t(d) * cap + r], e = pend[uint64_t(d) * cap + r];
        ok = ok && (s < o[d] + g.tile[d]) && (e > o[d]);
    }
    return ok;
}

// one CTA per tile: count (fill == false) or write (fill == true) the ids of the regions whose pixel box meets the tile,
// in ascending region order (ordered stream compaction: ballot + popc inside a warp, shared-memory scan across warps)
template<bool FILL>
__global__ void __launch_bounds__(256) tile_lists_kernel(TileGeom g, uint64_t nregions, uint64_t cap, uint64_t begin, uint64_t end,
                                                         const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ pend,
                                                         unsigned long long* __restrict__ counts, const uint64_t* __restrict__ offsets, uint32_t* __restrict__ list) {
    __shared__ uint32_t s_warp[8];
    __shared__ unsigned long long s_running;
    const uint64_t t = blockIdx.x;
    uint32_t o[3]; tile_origin(g, t, o);
    if (!tile_in_shard(g, o, begin, end)) { if (!FILL && threadIdx.x == 0) counts[t] = 0; return; }
    if (threadIdx.x == 0) s_running = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t base_out = FILL ? offsets[t] : 0;
    for (uint64_t base = 0; base < nregions; base += 256) {
        const uint64_t r = base + threadIdx.x;
        const bool hit = r < nregions && overlaps(g, o, pstart, pend, cap, r);
        const uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (uint32_t w = 0; w < 8; ++w) { const uint32_t c = s_warp[w]; if (w < warp) before += c; total += c; }
        const unsigned long long running = s_running;
        if (FILL && hit) list[base_out + running + before + __popc(m & ((1u << lane) - 1u))] = uint32_t(r);
        __syncthreads();
        if (threadIdx.x == 0) s_running = running + total;
        __syncthreads();
    }
    if (!FILL && threadIdx.x == 0) counts[t] = s_running;
}

// ---- region-major binning (used when regions x tiles is too large to test every pair) -------------------------------------
// tiles a region's pixel box touches, restricted to tiles that intersect the shard
__device__ __forceinline__ void region_tile_box(const TileGeom& g, const uint32_t* pstart, const uint32_t* pend, uint64_t cap, uint64_t r, uint32_t (&t0)[3], uint32_t (&t1)[3]) {
    for (int d = 0; d < 3; ++d) { t0[d] = 0; t1[d] = 1; }
    for (int d = 0; d < g.db; ++d) {
        const uint32_t s = pstart[uint64_t(d) * cap + r], e = pend[uint64_t(d) * cap + r];
        t0[d] = s / g.tile[d]; t1[d] = min((e + g.tile[d] - 1) / g.tile[d], g.tiles[d]);
    }
}
template<bool FILL>
__global__ void region_major_kernel(TileGeom g, uint64_t nregions, uint64_t cap, uint64_t begin, uint64_t end,
                                    const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ pend,
                                    unsigned long long* __restrict__ counts, const uint64_t* __restrict__ offsets, unsigned long long* __restrict__ cursor, uint32_t* __restrict__ list) {
    const uint64_t r = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= nregions) return;
    uint32_t t0[3], t1[3]; region_tile_box(g, pstart, pend, cap, r, t0, t1);
    for (uint32_t z = t0[2]; z < t1[2]; ++z) for (uint32_t y = t0[1]; y < t1[1]; ++y) for (uint32_t x = t0[0]; x < t1[0]; ++x) {
        const uint32_t o[3] = {x * g.tile[0], y * g.tile[1], z * g.tile[2]};
        if (!tile_in_shard(g, o, begin, end)) continue;
        const uint64_t t = uint64_t(x) + uint64_t(g.tiles[0]) * (uint64_t(y) + uint64_t(g.tiles[1]) * z);
        if (!FILL) atomicAdd(&counts[t], 1ull);
        else list[offsets[t] + atomicAdd(&cursor[t], 1ull)] = uint32_t(r);
    }
}
// restore table order inside every tile list: bitonic sort in shared memory when the padded list fits (<= smem_limit ids, 32768 in
// production), otherwise in place in global memory (rare: one tile touched by more than 32768 regions).  The global path cannot
// materialise the padding up to a power of two, so it runs the ALL-ASCENDING form of the network — the first step of every merge
// compares i with i ^ (k-1), the following ones i with i ^ j, every comparator orders (low index, high index) ascending — for which
// virtual +inf entries behind the end are never moved: a comparator that reaches past `len` is a no-op.  (The alternating-direction
// form used in shared memory would have to move +inf DOWN in its descending blocks; round 1 shipped that form here and lost ids
// for every non-power-of-two length)
__global__ void __launch_bounds__(1024) tile_sort_kernel(const uint64_t* __restrict__ offsets, uint32_t* __restrict__ list, uint32_t smem_limit) {
    extern __shared__ uint32_t s_ids[];
    const uint64_t lo = offsets[blockIdx.x], len = offsets[blockIdx.x + 1] - lo;
    if (len < 2) return;
    uint64_t P = 1; while (P < len) P <<= 1;
    uint32_t* a = list + lo;
    if (P <= smem_limit) {
        for (uint64_t i = threadIdx.x; i < P; i += blockDim.x) s_ids[i] = i < len ? a[i] : 0xffffffffu;
        __syncthreads();
        const uint32_t P32 = uint32_t(P), half = P32 >> 1;
        for (uint32_t k = 2; k <= P32; k <<= 1) for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t p = threadIdx.x; p < half; p += blockDim.x) {       // one thread per comparator: i has bit j clear, its partner has it set
                const uint32_t i = ((p & ~(j - 1u)) << 1) | (p & (j - 1u)), l = i | j;
                const bool up = (i & k) == 0;
                const uint32_t x = s_ids[i], y = s_ids[l];
                if ((x > y) == up) { s_ids[i] = y; s_ids[l] = x; }
            }
            __syncthreads();
        }
        for (uint64_t i = threadIdx.x; i < len; i += blockDim.x) a[i] = s_ids[i];
        return;
    }
    for (uint64_t k = 2; k <= P; k <<= 1) for (uint64_t j = k >> 1; j > 0; j >>= 1) {
        const uint64_t mask = (j == (k >> 1)) ? (k - 1) : j;
        for (uint64_t i = threadIdx.x; i < len; i += blockDim.x) {
            const uint64_t l = i ^ mask;
            if (l > i && l < len) {
                const uint32_t x = a[i], y = a[l];
                if (x > y) { a[i] = y; a[l] = x; }
            }
        }
        __syncthreads();
    }
}

// closed-form integral of the region's (marginalised) tensor-product interpolant over bin ∩ region:
// Region::integral_subrange -> sub_last (region.h:141-169), folding subrange(a_d,b_d) over the binned dims DB-1, ..., 0.
template<int S, int DB, class T>
__device__ __forceinline__ T patch_subrange(const T* patch, const T (&na)[3], const T (&nb)[3]) {
    if (DB == 1) return R::subrange<S, T>(na[0], nb[0], patch);
    if (DB == 2) {
        T t[S];
#pragma unroll
        for (int i0 = 0; i0 < S; ++i0) {
            T line[S];
#pragma unroll
            for (int i1 = 0; i1 < S; ++i1) line[i1] = patch[i0 + S * i1];
            t[i0] = R::subrange<S, T>(na[1], nb[1], line);
        }
        return R::subrange<S, T>(na[0], nb[0], t);
    }
    T t1[S];
#pragma unroll
    for (int i0 = 0; i0 < S; ++i0) {
        T t2[S];
#pragma unroll
        for (int i1 = 0; i1 < S; ++i1) {
            T line[S];
#pragma unroll
            for (int i2 = 0; i2 < S; ++i2) line[i2] = patch[i0 + S * i1 + S * S * i2];
            t2[i1] = R::subrange<S, T>(na[2], nb[2], line);
        }
        t1[i0] = R::subrange<S, T>(na[1], nb[1], t2);
    }
    return R::subrange<S, T>(na[0], nb[0], t1);
}

// Region::pdf_integral_subrange -> pdf_sub over the binned dims of a pdf patch (Simpson), same fold order as patch_subrange
template<int DB>
__device__ __forceinline__ float pdf_patch_subrange(const float* patch, const float (&na)[3], const float (&nb)[3]) {
    constexpr int S = 3;
    if (DB == 1) return R::simpson_pdf_integral_subrange<float>(na[0], nb[0], patch);
    if (DB == 2) {
        float t[S];
#pragma unroll
        for (int i0 = 0; i0 < S; ++i0) {
            float line[S];
#pragma unroll
            for (int i1 = 0; i1 < S; ++i1) line[i1] = patch[i0 + S * i1];
            t[i0] = R::simpson_pdf_integral_subrange<float>(na[1], nb[1], line);
        }
        return R::simpson_pdf_integral_subrange<float>(na[0], nb[0], t);
    }
    float t1[S];
#pragma unroll
    for (int i0 = 0; i0 < S; ++i0) {
        float t2[S];
#pragma unroll
        for (int i1 = 0; i1 < S; ++i1) {
            float line[S];
#pragma unroll
            for (int i2 = 0; i2 < S; ++i2) line[i2] = patch[i0 + S * i1 + S * S * i2];
            t2[i1] = R::simpson_pdf_integral_subrange<float>(na[2], nb[2], line);
        }
        t1[i0] = R::simpson_pdf_integral_subrange<float>(na[1], nb[1], t2);
    }
    return R::simpson_pdf_integral_subrange<float>(na[0], nb[0], t1);
}

template<int S, int DB, class T> struct Staged {
    static constexpr int P = (DB == 1 ? S : DB == 2 ? S * S : S * S * S);
    T patch[P]; T rmin[DB], rmax[DB]; T volume; uint32_t ps[DB], pe[DB];
};

// integral of region `rg` over (bin at `pos`) ∩ region, exactly as the reference evaluates it; false = empty intersection (skipped upstream)
template<int S, int DB, class T, class RG>
__device__ __forceinline__ bool pair_integral(const DomT<T>& dom, const uint32_t (&pos)[3], const RG& rg, T* integral) {
    // Range::intersection (range.h:92-101) in the binned dims; the other dims are the region's own extent
    T na[3], nb[3]; bool empty = false;
#pragma unroll
    for (int d = 0; d < DB; ++d) {       // bin box: min + Float(pos)*drange (regions-integrator-sequential.h:42-51)
        const T lo = R::add(dom.rmin[d], R::mul(T(pos[d]), dom.drange[d]));
        const T hi = R::add(dom.rmin[d], R::mul(T(pos[d] + 1u), dom.drange[d]));
        const T a = R::maxv(lo, rg.rmin[d]);
        const T b = R::maxv(a, R::minv(hi, rg.rmax[d]));
        empty = empty || (a >= b);
        na[d] = R::pos_in_range<T>(rg.rmin[d], rg.rmax[d], a);
        nb[d] = R::pos_in_range<T>(rg.rmin[d], rg.rmax[d], b);
    }
    *integral = R::mul(rg.volume, patch_subrange<S, DB, T>(rg.patch, na, nb));
    return !empty;
}

// Pre-pass of the SMALL path: one thread per (tile-list entry, k) over the WHOLE grid integrates the k-th bin a small region covers
// inside its tile (at most KS = 4 bins), so the slivers that crowd one tile along a discontinuity (20 000 in a single 16x16 tile of the
// tolerance benchmark) are spread over all SMs instead of being one CTA's serial work.  contrib[e*4+k] / ok[e*4+k]; area[e] = 0 marks
// an entry whose region is not small there (the bin threads integrate it themselves).
constexpr int WALK_KS = 4;
template<int S, int DB, class T>
__global__ void __launch_bounds__(256) walk_small_pairs_kernel(TileGeom g, DomT<T> dom, uint64_t cap, uint64_t ntiles, uint64_t nentries,
                                                               const T* __restrict__ patches, const T* __restrict__ rmin, const T* __restrict__ rmax,
                                                               const T* __restrict__ volume, const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ pend,
                                                               const uint64_t* __restrict__ offsets, const uint32_t* __restrict__ list,
                                                               T* __restrict__ contrib, unsigned char* __restrict__ ok, unsigned char* __restrict__ area_out) {
    const uint64_t item = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t e = item / WALK_KS; const int k = int(item % WALK_KS);
    if (e >= nentries) return;
    // tile of this entry: last t with offsets[t] <= e
    uint64_t lo = 0, hi = ntiles;
    while (hi - lo > 1) { const uint64_t mid = (lo + hi) >> 1; if (offsets[mid] <= e) lo = mid; else hi = mid; }
    uint32_t o[3]; tile_origin(g, lo, o);
    const uint64_t r = list[e];
    uint32_t blo[DB], bw[DB]; uint32_t area = 1;
    Staged<S, DB, T> rg;
#pragma unroll
    for (int d = 0; d < DB; ++d) {
        const uint32_t ps = pstart[uint64_t(d) * cap + r], pe = pend[uint64_t(d) * cap + r];
        const uint32_t a0 = max(ps, o[d]), a1 = min(min(pe, o[d] + g.tile[d]), g.res[d]);
        blo[d] = a0; bw[d] = a1 > a0 ? a1 - a0 : 0u;
        area = (bw[d] == 0u || area > uint32_t(WALK_KS)) ? (bw[d] == 0u ? 0u : uint32_t(WALK_KS) + 1u) : area * bw[d];
    }
    const bool small = area >= 1u && area <= uint32_t(WALK_KS);
    if (k == 0) area_out[e] = (unsigned char)(small ? area : 0u);
    if (!small || uint32_t(k) >= area) return;
#pragma unroll
    for (int d = 0; d < DB; ++d) { rg.rmin[d] = rmin[uint64_t(d) * cap + r]; rg.rmax[d] = rmax[uint64_t(d) * cap + r]; }
#pragma unroll
    for (int q = 0; q < Staged<S, DB, T>::P; ++q) rg.patch[q] = patches[uint64_t(q) * cap + r];
    rg.volume = volume[r];
    uint32_t p[3] = {0, 0, 0}; uint32_t kk = uint32_t(k);
#pragma unroll
    for (int d = 0; d < DB; ++d) { p[d] = blo[d] + kk % bw[d]; kk /= bw[d]; }
    T v; const bool good = pair_integral<S, DB, T>(dom, p, rg, &v);
    contrib[item] = v; ok[item] = good ? 1 : 0;
}

// K8/K10: one CTA per bin tile, one thread per bin.  The tile's region list is staged through shared memory in chunks
// (patches + boxes: the "region tree" a bin needs), every thread walks the chunk in table order, keeps the regions whose
// pixel box contains its bin and accumulates nbins * integral_subrange(bin ∩ region) with the reference's promotions:
//   bins(pos) += double(factor) * float        (regions-integrator-sequential.h:54; ...-variance-reduction.h:80)
// Small regions (<= KS bins of this tile: the slivers a tolerance-driven refinement piles up along a discontinuity) are integrated by a
// thread per (region, bin) pair in a grid-wide pre-pass (walk_small_pairs_kernel) and the bin threads only pick the values up in table
// order — otherwise a chunk of 64 slivers that all touch the same few bins is 64 integrals evaluated one after the other by a single
// lane (measured: 88 ms -> see profiles/results_r1.md for an 841 644-leaf table at 512x512 bins).  The summation order per bin, hence
// every bit, is unchanged.
// SMALL = false compiles that path out (tables of few large regions, e.g. BASELINE config 4: fewer registers, more resident CTAs).
template<int S, int DB, class T, bool SMALL>
__global__ void __launch_bounds__(256) walk_accumulate_kernel(TileGeom g, DomT<T> dom, uint64_t cap, uint64_t begin, uint64_t end, uint64_t nbins_total,
                                                              const T* __restrict__ patches, const T* __restrict__ rmin, const T* __restrict__ rmax,
                                                              const T* __restrict__ volume, const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ pend,
                                                              const uint64_t* __restrict__ offsets, const uint32_t* __restrict__ list,
                                                              int mode, T* __restrict__ out, T* __restrict__ approx, uint32_t* __restrict__ count,
                                                              const T* __restrict__ g_contrib, const unsigned char* __restrict__ g_ok, const unsigned char* __restrict__ g_area) {
    using St = Staged<S, DB, T>;
    constexpr int CHUNK = (St::P * sizeof(T) > 256) ? 16 : 64;
    constexpr int KS = WALK_KS;
    __shared__ St s_reg[CHUNK];
    __shared__ T s_contrib[CHUNK][KS];
    __shared__ unsigned char s_area[CHUNK];            // bins of this tile the region covers if that is 1..KS, else 0 (bin threads integrate it themselves)
    __shared__ unsigned char s_ok[CHUNK][KS];          // 0 = empty intersection
    __shared__ uint32_t s_lo[CHUNK][DB], s_w[CHUNK][DB];
    // LARGE regions over a 2-D bin grid (!SMALL): the first fold of integral_subrange — along bin dimension 1 — only depends on the bin's
    // ROW, so the 16 bins of a tile row would repeat it; it is evaluated once per (region, row, line) by the whole CTA and shared
    // (3.4x fewer line integrals at S = 3; same operations on the same operands, so the same bits).
    constexpr bool ROWS = !SMALL && DB == 2 && sizeof(T) == 4;
    constexpr int TR = 16;
    __shared__ T s_t[ROWS ? CHUNK : 1][ROWS ? TR : 1][ROWS ? S : 1];
    __shared__ unsigned char s_e1[ROWS ? CHUNK : 1][ROWS ? TR : 1];
    const uint64_t t = blockIdx.x;
    uint32_t o[3]; tile_origin(g, t, o);
    if (!tile_in_shard(g, o, begin, end)) return;
    // this thread's bin
    uint32_t pos[3] = {0, 0, 0}; { uint32_t k = threadIdx.x; for (int d = 0; d < DB; ++d) { pos[d] = o[d] + k % g.tile[d]; k /= g.tile[d]; } }
    bool live = true; uint64_t bin = 0, prod = 1;
    for (int d = 0; d < DB; ++d) { live = live && pos[d] < g.res[d]; bin += uint64_t(pos[d]) * prod; prod *= g.res[d]; }
    live = live && bin >= begin && bin < end;
    T acc = (mode == 0 && live) ? out[bin] : T(0);
    uint32_t cnt = 0;
    const double factor = double(nbins_total);
    const uint64_t lo = offsets[t], hi = offsets[t + 1];
    for (uint64_t base = lo; base < hi; base += CHUNK) {
        const int n = int(min(uint64_t(CHUNK), hi - base));
        __syncthreads();
        for (int k = threadIdx.x; k < n * St::P; k += blockDim.x) {
            const int j = k / St::P, q = k % St::P;
            if (SMALL && g_area[base + j] != 0) continue;       // integrated by the pre-pass: nobody reads its patch here
            s_reg[j].patch[q] = patches[uint64_t(q) * cap + list[base + j]];
        }
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            const uint64_t r = list[base + j];
            for (int d = 0; d < DB; ++d) {
                s_reg[j].rmin[d] = rmin[uint64_t(d) * cap + r]; s_reg[j].rmax[d] = rmax[uint64_t(d) * cap + r];
                const uint32_t ps = pstart[uint64_t(d) * cap + r], pe = pend[uint64_t(d) * cap + r];
                s_reg[j].ps[d] = ps; s_reg[j].pe[d] = pe;
                // the part of the region's pixel box inside this tile (and inside the grid)
                const uint32_t a0 = max(ps, o[d]), a1 = min(min(pe, o[d] + g.tile[d]), g.res[d]);
                s_lo[j][d] = a0; s_w[j][d] = a1 > a0 ? a1 - a0 : 0u;
            }
            s_reg[j].volume = volume[r];
            s_area[j] = SMALL ? g_area[base + j] : (unsigned char)0;
        }
        if constexpr (SMALL) {       // the small regions' integrals come from the grid-wide pre-pass (walk_small_pairs_kernel)
            for (int item = threadIdx.x; item < n * KS; item += blockDim.x) {
                const int j = item / KS, k = item % KS;
                if (k < int(g_area[base + j])) { s_contrib[j][k] = g_contrib[(base + j) * KS + k]; s_ok[j][k] = g_ok[(base + j) * KS + k]; }
            }
        }
        __syncthreads();
        if constexpr (ROWS) {
            for (int item = threadIdx.x; item < n * TR * S; item += blockDim.x) {
                const int j = item / (TR * S), rem = item % (TR * S), ry = rem / S, i0 = rem % S;
                const St& rg = s_reg[j];
                const uint32_t p1 = o[1] + uint32_t(ry);
                const T lo1 = R::add(dom.rmin[1], R::mul(T(p1), dom.drange[1])), hi1 = R::add(dom.rmin[1], R::mul(T(p1 + 1u), dom.drange[1]));
                const T a = R::maxv(lo1, rg.rmin[1]), b = R::maxv(a, R::minv(hi1, rg.rmax[1]));
                T line[S];
#pragma unroll
                for (int i1 = 0; i1 < S; ++i1) line[i1] = rg.patch[i0 + S * i1];
                s_t[j][ry][i0] = R::subrange<S, T>(R::pos_in_range<T>(rg.rmin[1], rg.rmax[1], a), R::pos_in_range<T>(rg.rmin[1], rg.rmax[1], b), line);
                if (i0 == 0) s_e1[j][ry] = (a >= b) ? 1 : 0;
            }
            __syncthreads();
        }
        if (live) {
            for (int j = 0; j < n; ++j) {
                const St& rg = s_reg[j];
                bool inside = true;
#pragma unroll
                for (int d = 0; d < DB; ++d) inside = inside && pos[d] >= rg.ps[d] && pos[d] < rg.pe[d];
                if (!inside) continue;
                ++cnt;
                T integral; bool ok;
                if constexpr (ROWS) {
                    const int ry = int(pos[1] - o[1]);
                    const T lo0 = R::add(dom.rmin[0], R::mul(T(pos[0]), dom.drange[0])), hi0 = R::add(dom.rmin[0], R::mul(T(pos[0] + 1u), dom.drange[0]));
                    const T a = R::maxv(lo0, rg.rmin[0]), b = R::maxv(a, R::minv(hi0, rg.rmax[0]));
                    T tt[S];
#pragma unroll
                    for (int i0 = 0; i0 < S; ++i0) tt[i0] = s_t[j][ry][i0];
                    integral = R::mul(rg.volume, R::subrange<S, T>(R::pos_in_range<T>(rg.rmin[0], rg.rmax[0], a), R::pos_in_range<T>(rg.rmin[0], rg.rmax[0], b), tt));
                    ok = !((a >= b) || s_e1[j][ry] != 0);
                } else if (SMALL && s_area[j] != 0) {
                    uint32_t k = 0, stride = 1;
#pragma unroll
                    for (int d = 0; d < DB; ++d) { k += (pos[d] - s_lo[j][d]) * stride; stride *= s_w[j][d]; }
                    integral = s_contrib[j][k]; ok = s_ok[j][k] != 0;
                } else {
                    ok = pair_integral<S, DB, T>(dom, pos, rg, &integral);
                }
                if (!ok) continue;                                             // regions-integrator-sequential.h:54 `if (!empty())`
                acc = R::from_double<T>(R::da(double(acc), R::dm(factor, double(integral))));
            }
        }
    }
    if (live) {
        if (mode == 0) out[bin] = acc;
        else { approx[bin - begin] = acc; count[bin - begin] = cnt; }
    }
}

// ---- throughput form of the walk (control variates of a FAST integrand: statistical parity, no bit contract) ------------------------
// The same tile walk in plain fp32 with FMAs: the antiderivative of a line's interpolant is evaluated in float (the reference promotes it
// to double, which is what makes the exact kernel issue-bound on FP64 and float<->double conversions: 6.0 ms at BASELINE config 4), the
// reciprocal extents of a region are computed once when it is staged, and the line fold along bin dimension 1 leaves its MONOMIAL
// coefficients pre-divided for the antiderivative in shared memory, so a (bin, region) pair costs an inside test, two clamps, two
// 3-term Horner evaluations and one FMA into the running sum.  Per-chunk partial sums go into a double accumulator (one DADD per 16
// regions), so the control variate of a bin is good to ~1e-7 relative.  2-D bin grids, any S.
template<int S> struct FastLine { float c[S]; };      // c[k] = monomial coefficient k of the line, divided by k+1
template<int S> __device__ __forceinline__ void fast_coefficients(const float* p, float* c) {
    if constexpr (S == 2) { c[0] = p[0]; c[1] = p[1] - p[0]; }
    else if constexpr (S == 3) { c[0] = p[0]; c[1] = fmaf(4.0f, p[1], fmaf(-3.0f, p[0], -p[2])); c[2] = fmaf(-4.0f, p[1], 2.0f * (p[0] + p[2])); }
    else {
        constexpr float k3 = 1.0f / 3.0f;
        c[0] = p[0];
        c[1] = (-25.0f * p[0] + 48.0f * p[1] - 36.0f * p[2] + 16.0f * p[3] - 3.0f * p[4]) * k3;
        c[2] = (70.0f * p[0] - 208.0f * p[1] + 228.0f * p[2] - 112.0f * p[3] + 22.0f * p[4]) * k3;
        c[3] = (-80.0f * p[0] + 288.0f * p[1] - 384.0f * p[2] + 224.0f * p[3] - 48.0f * p[4]) * k3;
        c[4] = (32.0f * p[0] - 128.0f * p[1] + 192.0f * p[2] - 128.0f * p[3] + 32.0f * p[4]) * k3;
    }
}
// integral over [a,b] (normalised) of the polynomial with pre-divided coefficients q[k] = c[k]/(k+1): F(b) - F(a), F(x) = x * Horner(q, x)
template<int S> __device__ __forceinline__ float fast_subrange(float a, float b, const float* q) {
    float fa = q[S - 1], fb = q[S - 1];
#pragma unroll
    for (int k = S - 2; k >= 0; --k) { fa = fmaf(fa, a, q[k]); fb = fmaf(fb, b, q[k]); }
    return fmaf(fb, b, -fa * a);
}
template<int S>
__global__ void __launch_bounds__(256) walk_accumulate_fast_kernel(TileGeom g, DomT<float> dom, uint64_t cap, uint64_t begin, uint64_t end, uint64_t nbins_total,
                                                                   const float* __restrict__ patches, const float* __restrict__ rmin, const float* __restrict__ rmax,
                                                                   const float* __restrict__ volume, const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ pend,
                                                                   const uint64_t* __restrict__ offsets, const uint32_t* __restrict__ list,
                                                                   int mode, float* __restrict__ out, float* __restrict__ approx, uint32_t* __restrict__ count) {
    // The integral of a region's S x S patch over bin (x, y) of the tile separates: sum_k q[region][y][k] * (b^(k+1) - a^(k+1)), where q are the
    // pre-divided coefficients of the patch folded along bin dimension 1 over row y (they depend on (region, row) only) and [a, b] is the bin's
    // normalised extent along dimension 0 (it depends on (region, column) only).  Both tables and the region's pixel box — as a row mask and
    // a column mask — are built once per chunk of regions by the whole CTA; a (bin, region) pair is then a mask test and an S-term dot product.
    constexpr int CHUNK = 32, TR = 16, P = S * S, SP = S <= 4 ? 4 : 8;
    struct Reg { float patch[P]; float rmin[2], rmax[2], inv[2]; float scale; uint32_t ps[2], pe[2]; };
    __shared__ Reg s_reg[CHUNK];
    __shared__ __align__(16) float s_q[CHUNK][TR][SP];      // per (region, tile row): coefficients c[k] / (k+1) * volume * nbins; zero where the row misses the region
    __shared__ __align__(16) float s_p[CHUNK][TR][SP];      // per (region, tile column): b^(k+1) - a^(k+1); zero where the column misses the region
    __shared__ uint32_t s_mask[CHUNK];                      // pixel box: rows in the low half, columns in the high half
    const uint64_t t = blockIdx.x;
    uint32_t o[3]; tile_origin(g, t, o);
    if (!tile_in_shard(g, o, begin, end)) return;
    uint32_t pos[2]; pos[0] = o[0] + threadIdx.x % g.tile[0]; pos[1] = o[1] + threadIdx.x / g.tile[0];
    bool live = pos[0] < g.res[0] && pos[1] < g.res[1];
    const uint64_t bin = uint64_t(pos[0]) + uint64_t(pos[1]) * g.res[0];
    live = live && bin >= begin && bin < end;
    const int cx = int(pos[0] - o[0]), ry = int(pos[1] - o[1]);
    const uint32_t mybit = (1u << ry), mycol = (0x10000u << cx);
    double acc = (mode == 0 && live) ? double(out[bin]) : 0.0;
    uint32_t cnt = 0;
    const float factor = float(nbins_total);
    const uint64_t lo = offsets[t], hi = offsets[t + 1];
    for (uint64_t base = lo; base < hi; base += CHUNK) {
        const int n = int(min(uint64_t(CHUNK), hi - base));
        __syncthreads();
        for (int k = threadIdx.x; k < n * P; k += blockDim.x) { const int j = k / P, q = k % P; s_reg[j].patch[q] = patches[uint64_t(q) * cap + list[base + j]]; }
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            const uint64_t r = list[base + j];
            uint32_t m = 0;
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                const float a = rmin[uint64_t(d) * cap + r], b = rmax[uint64_t(d) * cap + r];
                s_reg[j].rmin[d] = a; s_reg[j].rmax[d] = b; s_reg[j].inv[d] = b > a ? 1.0f / (b - a) : 0.0f;
                const uint32_t ps = pstart[uint64_t(d) * cap + r], pe = pend[uint64_t(d) * cap + r];
                // bins [ps, pe) of the grid -> bits of the tile's 16 rows / columns
                const uint32_t l0 = ps > o[d] ? min(ps - o[d], 16u) : 0u, l1 = pe > o[d] ? min(pe - o[d], 16u) : 0u;
                const uint32_t bits = l1 > l0 ? (((1u << l1) - 1u) & ~((1u << l0) - 1u)) : 0u;
                m |= d == 0 ? (bits << 16) : bits;
            }
            s_reg[j].scale = volume[r] * factor;
            s_mask[j] = m;
        }
        __syncthreads();
        for (int item = threadIdx.x; item < n * TR; item += blockDim.x) {       // per (region, tile row): the fold along bin dimension 1; per (region, column): the powers
            const int j = item / TR, row = item % TR;
            const Reg& rg = s_reg[j];
            {
                const uint32_t p1 = o[1] + uint32_t(row);
                const float lo1 = fmaf(float(p1), dom.drange[1], dom.rmin[1]), hi1 = fmaf(float(p1 + 1u), dom.drange[1], dom.rmin[1]);
                const float a = fmaxf(lo1, rg.rmin[1]), b = fmaxf(a, fminf(hi1, rg.rmax[1]));
                const float na = (a - rg.rmin[1]) * rg.inv[1], nb = (b - rg.rmin[1]) * rg.inv[1];
                float tt[S];
#pragma unroll
                for (int i0 = 0; i0 < S; ++i0) {
                    float line[S], c[S];
#pragma unroll
                    for (int i1 = 0; i1 < S; ++i1) line[i1] = rg.patch[i0 + S * i1];
                    fast_coefficients<S>(line, c);
#pragma unroll
                    for (int k = 0; k < S; ++k) c[k] *= 1.0f / float(k + 1);
                    tt[i0] = fast_subrange<S>(na, nb, c);
                }
                float c[S]; fast_coefficients<S>(tt, c);
                const float sc = (a >= b) ? 0.0f : rg.scale;                    // empty intersection: skipped upstream (regions-integrator-sequential.h:54)
#pragma unroll
                for (int k = 0; k < SP; ++k) s_q[j][row][k] = k < S ? c[k] * (sc / float(k + 1)) : 0.0f;      // volume * nbins folded in
            }
            {
                const uint32_t p0 = o[0] + uint32_t(row);                       // the same index as a column
                const float lo0 = fmaf(float(p0), dom.drange[0], dom.rmin[0]), hi0 = fmaf(float(p0 + 1u), dom.drange[0], dom.rmin[0]);
                const float a = fmaxf(lo0, rg.rmin[0]), b = fmaxf(a, fminf(hi0, rg.rmax[0]));
                const float na = (a - rg.rmin[0]) * rg.inv[0], nb = (b - rg.rmin[0]) * rg.inv[0];
                float pa = na, pb = nb;
#pragma unroll
                for (int k = 0; k < SP; ++k) { s_p[j][row][k] = (k < S && a < b) ? pb - pa : 0.0f; pa *= na; pb *= nb; }
            }
        }
        __syncthreads();
        if (live) {
            float part = 0.0f;
#pragma unroll 4
            for (int j = 0; j < n; ++j) {
                const uint32_t m = s_mask[j];
                if (!((m & mybit) && (m & mycol))) continue;
                ++cnt;
                const float4 q0 = *reinterpret_cast<const float4*>(&s_q[j][ry][0]), p0 = *reinterpret_cast<const float4*>(&s_p[j][cx][0]);
                float v = q0.x * p0.x;
                v = fmaf(q0.y, p0.y, v);
                if constexpr (S >= 3) v = fmaf(q0.z, p0.z, v);
                if constexpr (S >= 4) v = fmaf(q0.w, p0.w, v);
                if constexpr (S >= 5) { v = fmaf(s_q[j][ry][4], s_p[j][cx][4], v); }
                part += v;
            }
            acc += double(part);
        }
    }
    if (live) {
        if (mode == 0) out[bin] = float(acc);
        else { approx[bin - begin] = float(acc); count[bin - begin] = cnt; }
    }
}

template<int S, int DB, class T>
int launch_accumulate(vb200_ctx* ctx, const vb200_regions* r, const BinWalkT<T>& w, const TileGeom& g, const DomT<T>& dom,
                      uint64_t begin, uint64_t end, uint64_t total, int mode, T* out, T* approx, uint32_t* count) {
    // regions that cover only a few bins each are integrated pair-parallel (SMALL); tables of few large regions keep the leaner kernel
    const bool small = r->count * 4ull >= total && w.pairs > 0;
    if (small) {
        T* contrib = nullptr; unsigned char* ok = nullptr; unsigned char* area = nullptr;
        const uint64_t items = w.pairs * uint64_t(WALK_KS);
        if (dmalloc(ctx, &contrib, items * sizeof(T)) != cudaSuccess || dmalloc(ctx, &ok, items) != cudaSuccess || dmalloc(ctx, &area, w.pairs) != cudaSuccess) {
            cudaGetLastError(); dfree(ctx, contrib); dfree(ctx, ok); dfree(ctx, area);
            return fail(ctx, VB200_ERR_NOMEM, "cudaMalloc failed (small-region contributions)");
        }
        walk_small_pairs_kernel<S, DB, T><<<unsigned((items + 255) / 256), 256, 0, ctx->stream>>>(g, dom, w.cap, w.ntiles, w.pairs, w.patches, RegCols<T>::rmin(r), RegCols<T>::rmax(r), w.volume,
                                                                                                   w.pstart, w.pend, w.tile_offset, w.tile_list, contrib, ok, area);
        walk_accumulate_kernel<S, DB, T, true><<<unsigned(w.ntiles), 256, 0, ctx->stream>>>(g, dom, w.cap, begin, end, total, w.patches, RegCols<T>::rmin(r), RegCols<T>::rmax(r), w.volume,
                                                                                   w.pstart, w.pend, w.tile_offset, w.tile_list, mode, out, approx, count, contrib, ok, area);
        ctx->launches += 2;
        const cudaError_t e = cudaGetLastError();
        dfree(ctx, contrib); dfree(ctx, ok); dfree(ctx, area);      // stream-ordered frees: after the kernels above
        VB200_CUDA(ctx, e);
        return VB200_OK;
    }
    walk_accumulate_kernel<S, DB, T, false><<<unsigned(w.ntiles), 256, 0, ctx->stream>>>(g, dom, w.cap, begin, end, total, w.patches, RegCols<T>::rmin(r), RegCols<T>::rmax(r), w.volume,
                                                                                   w.pstart, w.pend, w.tile_offset, w.tile_list, mode, out, approx, count, nullptr, nullptr, nullptr);
    ctx->launches++;
    VB200_CUDA(ctx, cudaGetLastError());
    return VB200_OK;
}

// ---- weighted Russian roulette among the regions of a bin -------------------------------------------------
//   rr_integral_region  region-russian-roulette.h:30-67   weight = |integral_subrange(bin ∩ region)|                       (policy 1)
//   rr_error_region     region-russian-roulette.h:69-106  weight = |region.error()| * vol(bin ∩ region) / vol(region)       (policy 2)
//   rr_pdf_region       region-russian-roulette.h:108-147 weight = pdf_integral_subrange(bin ∩ region) (Simpson; `patches` = pdf patches) (policy 3)
// then, per bin: sum; w' = sum <= 0 ? 1 : max(w, floor_factor*sum/n)  (floor_factor 0.01; double(0.01f) for rr_pdf_region); std::discrete_distribution over w' (probabilities w'/sum(w')).
// The weights are never materialised (1e9 pairs at BASELINE config 4): three more walks over the tile lists recompute them —
// pass 1 sum(w), pass 2 sum(w'), pass 3 the per-sample inverse-CDF pick — all in table order, all sums in double like upstream.
template<int S, int DB> struct StagedW {
    static constexpr int P = (DB == 1 ? S : DB == 2 ? S * S : S * S * S);
    float patch[P]; float rmin[DB], rmax[DB]; float volume; uint32_t ps[DB], pe[DB];
    float ext[VB200_MAX_DIM]; float rerr; uint32_t id;
};

template<int S, int DB, class RG>
__device__ __forceinline__ double pair_weight(int policy, int next, const float (&ba)[DB], const float (&bb)[DB], const RG& rg) {
    float na[3], nb[3]; float vol = 1.0f;
#pragma unroll
    for (int d = 0; d < DB; ++d) {           // Range::intersection (range.h:92-101)
        const float a = R::maxv(ba[d], rg.rmin[d]);
        const float b = R::maxv(a, R::minv(bb[d], rg.rmax[d]));
        na[d] = R::pos_in_range<float>(rg.rmin[d], rg.rmax[d], a);
        nb[d] = R::pos_in_range<float>(rg.rmin[d], rg.rmax[d], b);
        vol = R::fm(vol, R::fs(b, a));
    }
    if (policy == 1) return double(fabsf(R::fm(rg.volume, patch_subrange<S, DB, float>(rg.patch, na, nb))));       // :45, NormDefault = abs
    if (policy == 3) {                                                                                              // :125, the patch is the pdf patch
        if constexpr (S == 3) return double(R::fm(rg.volume, pdf_patch_subrange<DB>(rg.patch, na, nb)));
        else return 0.0;
    }
    for (int e = 0; e < next; ++e) vol = R::fm(vol, rg.ext[e]);                                                     // Range::volume, dimension order
    return double(R::fd(R::fm(fabsf(rg.rerr), vol), rg.volume));                                                    // :86 (float arithmetic)
}

template<int S, int DB>
__global__ void __launch_bounds__(256) walk_rr_kernel(TileGeom g, DomT<float> dom, uint64_t cap, uint64_t begin, uint64_t end, uint64_t base, int D, int policy, int pass, double floor_factor,
                                                      const float* __restrict__ patches, const float* __restrict__ rmin, const float* __restrict__ rmax,
                                                      const float* __restrict__ volume, const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ pend,
                                                      const uint64_t* __restrict__ offsets, const uint32_t* __restrict__ list, const float* __restrict__ rerr,
                                                      const uint32_t* __restrict__ count, double* __restrict__ wsum, double* __restrict__ csum,
                                                      uint32_t spp, const uint32_t* __restrict__ raw, uint32_t* __restrict__ chosen) {
    using St = StagedW<S, DB>;
    constexpr int CHUNK = (St::P * sizeof(float) > 256) ? 16 : 32;
    __shared__ St s_reg[CHUNK];
    const uint64_t t = blockIdx.x;
    uint32_t o[3]; tile_origin(g, t, o);
    if (!tile_in_shard(g, o, begin, end)) return;
    uint32_t pos[3] = {0, 0, 0}; { uint32_t k = threadIdx.x; for (int d = 0; d < DB; ++d) { pos[d] = o[d] + k % g.tile[d]; k /= g.tile[d]; } }
    bool live = true; uint64_t bin = 0, prod = 1;
    for (int d = 0; d < DB; ++d) { live = live && pos[d] < g.res[d]; bin += uint64_t(pos[d]) * prod; prod *= g.res[d]; }
    live = live && bin >= begin && bin < end;
    float ba[DB], bb[DB];
#pragma unroll
    for (int d = 0; d < DB; ++d) {
        ba[d] = R::add(dom.rmin[d], R::mul(float(pos[d]), dom.drange[d]));
        bb[d] = R::add(dom.rmin[d], R::mul(float(pos[d] + 1u), dom.drange[d]));
    }
    const int next = D - DB;
    const uint64_t nb = end - begin, b = bin - begin;
    double acc = 0.0, floor_w = 0.0, total = 0.0; bool flat = false;
    if (live && pass >= 2) {
        const double ws = wsum[bin - base];
        flat = ws <= 0.0;                                                                   // :49 / :88
        floor_w = R::dd(R::dm(floor_factor, ws), double(count[bin - base]));                // :50 / :89 / :130
        if (pass == 3) total = csum[bin - base];
    }
    uint32_t last_id = 0;
    const uint64_t lo = offsets[t], hi = offsets[t + 1];
    for (uint64_t cb = lo; cb < hi; cb += CHUNK) {
        const int n = int(min(uint64_t(CHUNK), hi - cb));
        __syncthreads();
        for (int k = threadIdx.x; k < n * St::P; k += blockDim.x) {
            const int j = k / St::P, q = k % St::P;
            s_reg[j].patch[q] = patches[uint64_t(q) * cap + list[cb + j]];
        }
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            const uint64_t r = list[cb + j];
            for (int d = 0; d < DB; ++d) {
                s_reg[j].rmin[d] = rmin[uint64_t(d) * cap + r]; s_reg[j].rmax[d] = rmax[uint64_t(d) * cap + r];
                s_reg[j].ps[d] = pstart[uint64_t(d) * cap + r]; s_reg[j].pe[d] = pend[uint64_t(d) * cap + r];
            }
            for (int e = 0; e < next; ++e) s_reg[j].ext[e] = R::fs(rmax[uint64_t(DB + e) * cap + r], rmin[uint64_t(DB + e) * cap + r]);
            s_reg[j].volume = volume[r]; s_reg[j].rerr = rerr ? rerr[r] : 0.0f; s_reg[j].id = uint32_t(r);
        }
        __syncthreads();
        if (!live) continue;
        double cum[CHUNK]; unsigned char which[CHUNK]; int k = 0;
        const double c0 = acc;
        for (int j = 0; j < n; ++j) {
            const St& rg = s_reg[j];
            bool inside = true;
#pragma unroll
            for (int d = 0; d < DB; ++d) inside = inside && pos[d] >= rg.ps[d] && pos[d] < rg.pe[d];
            if (!inside) continue;
            double w = pair_weight<S, DB>(policy, next, ba, bb, rg);
            if (pass >= 2) w = flat ? 1.0 : fmax(w, floor_w);
            acc = R::da(acc, w);
            if (pass == 3) { cum[k] = acc; which[k] = (unsigned char)j; ++k; last_id = rg.id; }
        }
        if (pass == 3 && k > 0) {
            for (uint32_t sidx = 0; sidx < spp; ++sidx) {
                const double tgt = R::dm(R::dm(double(raw[uint64_t(sidx) * nb + b]), 2.3283064365386963e-10), total);
                if (tgt >= c0 && tgt < acc) {
                    int i = 0; while (i < k - 1 && !(cum[i] > tgt)) ++i;
                    chosen[uint64_t(sidx) * nb + b] = s_reg[which[i]].id;
                }
            }
        }
    }
    if (!live) return;
    if (pass == 1) wsum[bin - base] = acc;
    else if (pass == 2) csum[bin - base] = acc;
    else for (uint32_t sidx = 0; sidx < spp; ++sidx) {       // u * total rounded up to the total itself: the last region
        const double tgt = R::dm(R::dm(double(raw[uint64_t(sidx) * nb + b]), 2.3283064365386963e-10), total);
        if (!(tgt < acc)) chosen[uint64_t(sidx) * nb + b] = last_id;
    }
}

// rrfactor = 1.0 / probabilities()[choice] of every residual sample (region-russian-roulette.h:59 / :98), probabilities = w'/sum(w')
// (libstdc++ discrete_distribution: bits/random.tcc:2657-2678; fewer than two weights -> {1.0})
template<int S, int DB>
__global__ void __launch_bounds__(128) rr_factor_kernel(DomT<float> dom, uint64_t cap, uint64_t s0, uint64_t nb, uint64_t base, int D, int policy, double floor_factor, uint32_t spp,
                                                        const float* __restrict__ patches, const float* __restrict__ rmin, const float* __restrict__ rmax,
                                                        const float* __restrict__ volume, const float* __restrict__ rerr, const uint32_t* __restrict__ count,
                                                        const double* __restrict__ wsum, const double* __restrict__ csum, const uint32_t* __restrict__ chosen,
                                                        double* __restrict__ rrf) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nb * spp) return;
    const uint64_t b = i % nb, bin = s0 + b;
    const uint32_t n = count[bin - base];
    if (n < 2) { rrf[i] = 1.0; return; }
    const uint64_t r = chosen[i];
    uint32_t pos[3]; { uint64_t q = bin; for (int d = 0; d < DB; ++d) { pos[d] = uint32_t(q % dom.res[d]); q /= dom.res[d]; } }
    StagedW<S, DB> rg; float ba[DB], bb[DB];
#pragma unroll
    for (int d = 0; d < DB; ++d) {
        ba[d] = R::add(dom.rmin[d], R::mul(float(pos[d]), dom.drange[d]));
        bb[d] = R::add(dom.rmin[d], R::mul(float(pos[d] + 1u), dom.drange[d]));
        rg.rmin[d] = rmin[uint64_t(d) * cap + r]; rg.rmax[d] = rmax[uint64_t(d) * cap + r];
    }
#pragma unroll
    for (int q = 0; q < StagedW<S, DB>::P; ++q) rg.patch[q] = patches[uint64_t(q) * cap + r];
    const int next = D - DB;
    for (int e = 0; e < next; ++e) rg.ext[e] = R::fs(rmax[uint64_t(DB + e) * cap + r], rmin[uint64_t(DB + e) * cap + r]);
    rg.volume = volume[r]; rg.rerr = rerr ? rerr[r] : 0.0f;
    const double ws = wsum[bin - base];
    double w = pair_weight<S, DB>(policy, next, ba, bb, rg);
    w = (ws <= 0.0) ? 1.0 : fmax(w, R::dd(R::dm(floor_factor, ws), double(n)));
    rrf[i] = R::dd(1.0, R::dd(w, csum[bin - base]));
}

// Region::error() (region.h:420-424): volume * (fold_all(high rule) - fold_all(low rule)); one thread per region, folds of
// dimension 0 in place (fold.h:87-108)
template<int SH, int SL>
__global__ void __launch_bounds__(64) region_total_error_kernel(uint64_t n, uint64_t cap, int sd, const float* __restrict__ data, const float* __restrict__ volume, float* __restrict__ rerr) {
    const uint64_t r = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float v[729];
    float res[2];
    for (int which = 0; which < 2; ++which) {
        for (int k = 0; k < sd; ++k) v[k] = data[uint64_t(k) * cap + r];
        for (int m = sd / SH; m >= 1; m /= SH) {
            for (int o = 0; o < m; ++o) {
                float line[SH];
#pragma unroll
                for (int e = 0; e < SH; ++e) line[e] = v[o * SH + e];
                v[o] = which == 0 ? R::apply<SH, float>(line) : R::low<SH, SL, float>(line);
            }
            if (m == 1) break;
        }
        res[which] = v[0];
    }
    rerr[r] = R::fm(volume[r], R::fs(res[0], res[1]));
}

template<class T> TileGeom make_geom(const BinWalkT<T>& w, const DomT<T>& dom) {
    TileGeom g; g.db = w.db;
    for (int d = 0; d < 3; ++d) { g.tile[d] = w.tile[d]; g.tiles[d] = w.tiles[d]; g.res[d] = d < w.db ? uint32_t(dom.res[d]) : 1u; }
    return g;
}

} // namespace

namespace vb200 {

template<class T> void walk_free_t(BinWalkT<T>* w) {
    vb200_ctx* ctx = w->ctx;
    if (!ctx) { *w = BinWalkT<T>(); return; }
    dfree(ctx, w->patches); dfree(ctx, w->volume); dfree(ctx, w->pstart); dfree(ctx, w->pend); dfree(ctx, w->tile_offset); dfree(ctx, w->tile_list);
    dfree(ctx, w->scratch[0]); dfree(ctx, w->scratch[1]);
    *w = BinWalkT<T>();
}

template<class T> int walk_build_t(vb200_ctx* ctx, const vb200_regions* r, const DomT<T>& dom, uint64_t begin, uint64_t end, BinWalkT<T>* w) {
    *w = BinWalkT<T>();
    w->ctx = ctx;
    const int S = r->SH, D = r->dim, db = dom.dimbins;
    const uint64_t n = r->count, cap = r->capacity;
    w->S = S; w->db = db; w->nregions = n; w->cap = cap;
    int patch = 1; for (int i = 0; i < db; ++i) patch *= S;
    w->patch = patch;
    auto bail = [&] (int code) { walk_free_t<T>(w); return code; };
#define VB200_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return bail(fail(ctx, VB200_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__))); } while (0)
    // 1. marginalise the non-binned dimensions D-1 .. db (each fold rounds to float, exactly like the lazy reference folds)
    const T* cur = RegCols<T>::data(r);
    if (D > db) {
        uint64_t biggest = 1; for (int i = 0; i < D - 1; ++i) biggest *= uint64_t(S);
        VB200_TRY(dmalloc(ctx, &w->scratch[0], biggest * cap * sizeof(T)));
        if (D - db > 1) VB200_TRY(dmalloc(ctx, &w->scratch[1], (biggest / uint64_t(S)) * cap * sizeof(T)));
        int which = 0;
        for (int m = D; m > db; --m) {
            int lower = 1; for (int i = 0; i < m - 1; ++i) lower *= S;
            T* dst = w->scratch[which];
            dim3 grid(unsigned((n + 127) / 128), unsigned(lower));
            if (S == 2) fold_last_dim_kernel<2, T><<<grid, 128, 0, ctx->stream>>>(n, cap, lower, cur, dst);
            else if (S == 3) fold_last_dim_kernel<3, T><<<grid, 128, 0, ctx->stream>>>(n, cap, lower, cur, dst);
            else fold_last_dim_kernel<5, T><<<grid, 128, 0, ctx->stream>>>(n, cap, lower, cur, dst);
            ctx->launches++;
            VB200_TRY(cudaGetLastError());
            cur = dst; which ^= 1;
        }
    }
    VB200_TRY(dmalloc(ctx, &w->patches, uint64_t(patch) * cap * sizeof(T)));
    VB200_TRY(cudaMemcpyAsync(w->patches, cur, uint64_t(patch) * cap * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    // 2. volumes + pixel boxes
    VB200_TRY(dmalloc(ctx, &w->volume, cap * sizeof(T)));
    VB200_TRY(dmalloc(ctx, &w->pstart, uint64_t(db) * cap * sizeof(uint32_t)));
    VB200_TRY(dmalloc(ctx, &w->pend, uint64_t(db) * cap * sizeof(uint32_t)));
    region_boxes_kernel<T><<<unsigned((n + 255) / 256), 256, 0, ctx->stream>>>(n, cap, D, db, dom, RegCols<T>::rmin(r), RegCols<T>::rmax(r), w->volume, w->pstart, w->pend);
    ctx->launches++;
    VB200_TRY(cudaGetLastError());
    // 3. tiles of 256 bins and their ordered region lists
    if (db == 1) { w->tile[0] = 256; } else if (db == 2) { w->tile[0] = 16; w->tile[1] = 16; } else { w->tile[0] = 8; w->tile[1] = 8; w->tile[2] = 4; }
    w->ntiles = 1;
    for (int d = 0; d < 3; ++d) { w->tiles[d] = d < db ? uint32_t((dom.res[d] + w->tile[d] - 1) / w->tile[d]) : 1u; w->ntiles *= w->tiles[d]; }
    if (w->ntiles > 0x7fffffffull) return bail(fail(ctx, VB200_ERR_UNSUPPORTED, "bin grid too large for the tile walk"));
    const TileGeom g = make_geom<T>(*w, dom);
    VB200_TRY(dmalloc(ctx, &w->tile_offset, (w->ntiles + 1) * sizeof(uint64_t)));
    unsigned long long* counts = reinterpret_cast<unsigned long long*>(w->tile_offset);
    // every (tile, region) pair tested by brute force keeps table order for free; beyond ~2.7e8 pairs bin the regions into
    // their tiles with atomics instead and sort each tile list back into table order
    uint64_t pair_limit = 1ull << 28;
    if (const char* env = std::getenv("VB200_TILE_PAIR_LIMIT")) pair_limit = std::strtoull(env, nullptr, 10);     // test knob
    const bool region_major = w->ntiles * n > pair_limit && w->ntiles > 1;
    if (region_major) {
        VB200_TRY(cudaMemsetAsync(counts, 0, (w->ntiles + 1) * sizeof(uint64_t), ctx->stream));
        region_major_kernel<false><<<unsigned((n + 255) / 256), 256, 0, ctx->stream>>>(g, n, cap, begin, end, w->pstart, w->pend, counts, nullptr, nullptr, nullptr);
    } else {
        tile_lists_kernel<false><<<unsigned(w->ntiles), 256, 0, ctx->stream>>>(g, n, cap, begin, end, w->pstart, w->pend, counts, nullptr, nullptr);
    }
    ctx->launches++;
    VB200_TRY(cudaGetLastError());
    std::vector<uint64_t> h(w->ntiles + 1);
    VB200_TRY(cudaMemcpyAsync(h.data(), counts, w->ntiles * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    VB200_TRY(cudaStreamSynchronize(ctx->stream));
    uint64_t run = 0; w->max_list = 0;
    for (uint64_t t = 0; t < w->ntiles; ++t) { const uint64_t c = h[t]; if (c > w->max_list) w->max_list = c; h[t] = run; run += c; } h[w->ntiles] = run;
    w->pairs = run;
    VB200_TRY(cudaMemcpyAsync(w->tile_offset, h.data(), (w->ntiles + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    VB200_TRY(dmalloc(ctx, &w->tile_list, (run + 1) * sizeof(uint32_t)));
    if (region_major) {
        unsigned long long* cursor = nullptr;
        VB200_TRY(dmalloc(ctx, &cursor, w->ntiles * sizeof(unsigned long long)));
        cudaError_t e1 = cudaMemsetAsync(cursor, 0, w->ntiles * sizeof(unsigned long long), ctx->stream);
        region_major_kernel<true><<<unsigned((n + 255) / 256), 256, 0, ctx->stream>>>(g, n, cap, begin, end, w->pstart, w->pend, nullptr, w->tile_offset, cursor, w->tile_list);
        cudaFuncSetAttribute(tile_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4);
        uint32_t smem_limit = 32768;        // VB200_TILE_SORT_SMEM_LIMIT: test knob that sends shorter lists down the global-memory path
        if (const char* env = std::getenv("VB200_TILE_SORT_SMEM_LIMIT")) { const long v = std::atol(env); if (v >= 0 && v < 32768) smem_limit = uint32_t(v); }
        // shared memory for the longest padded list only (the lengths are on the host): two CTAs per SM instead of one for lists of a few thousand ids
        uint64_t pmax = 1; while (pmax < w->max_list) pmax <<= 1;
        const size_t sort_smem = size_t(std::min<uint64_t>(pmax, smem_limit)) * 4;
        tile_sort_kernel<<<unsigned(w->ntiles), 1024, sort_smem, ctx->stream>>>(w->tile_offset, w->tile_list, smem_limit);
        ctx->launches += 2;
        cudaError_t e2 = cudaGetLastError(), e3 = cudaStreamSynchronize(ctx->stream);
        dfree(ctx, cursor);
        if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) return bail(fail(ctx, VB200_ERR_CUDA, "tile list construction failed: %s",
                cudaGetErrorString(e3 != cudaSuccess ? e3 : e2 != cudaSuccess ? e2 : e1)));
    } else {
        tile_lists_kernel<true><<<unsigned(w->ntiles), 256, 0, ctx->stream>>>(g, n, cap, begin, end, w->pstart, w->pend, nullptr, w->tile_offset, w->tile_list);
        ctx->launches++;
        VB200_TRY(cudaGetLastError());
        VB200_TRY(cudaStreamSynchronize(ctx->stream));       // h goes out of scope
    }
#undef VB200_TRY
    return VB200_OK;
}

template<class T> int walk_accumulate_t(vb200_ctx* ctx, const vb200_regions* r, const BinWalkT<T>& w, const DomT<T>& dom, uint64_t begin, uint64_t end,
                                        int mode, T* out, T* approx, uint32_t* count) {
    const TileGeom g = make_geom<T>(w, dom);
    const uint64_t total = nbins_of(dom);
#define VB200_WALK(SS, DD) if (w.S == SS && w.db == DD) return launch_accumulate<SS, DD, T>(ctx, r, w, g, dom, begin, end, total, mode, out, approx, count);
    VB200_WALK(2, 1) VB200_WALK(2, 2) VB200_WALK(2, 3) VB200_WALK(3, 1) VB200_WALK(3, 2) VB200_WALK(3, 3) VB200_WALK(5, 1) VB200_WALK(5, 2) VB200_WALK(5, 3)
#undef VB200_WALK
    return fail(ctx, VB200_ERR_UNSUPPORTED, "no bin walk for rule with %d samples and %d binned dimensions", w.S, w.db);
}

// throughput form (fp32 + FMA, no bit contract): 2-D bin grids of 16x16 tiles; false = shape not covered, the caller takes the exact walk
bool walk_accumulate_fast(vb200_ctx* ctx, const vb200_regions* r, const BinWalk& w, const vb200_domain& dom_, uint64_t begin, uint64_t end,
                          int mode, float* out, float* approx, uint32_t* count, int* rc) {
    *rc = VB200_OK;
    if (w.db != 2 || w.tile[0] != 16 || w.tile[1] != 16 || (w.S != 2 && w.S != 3 && w.S != 5)) return false;
    const DomT<float> dom = to_dom(dom_);
    const TileGeom g = make_geom<float>(w, dom);
    const uint64_t total = nbins_of(dom);
#define VB200_FASTWALK(SS) if (w.S == SS) walk_accumulate_fast_kernel<SS><<<unsigned(w.ntiles), 256, 0, ctx->stream>>>(g, dom, w.cap, begin, end, total, w.patches, r->rmin, r->rmax, w.volume, \
                                                                                                        w.pstart, w.pend, w.tile_offset, w.tile_list, mode, out, approx, count);
    VB200_FASTWALK(2) VB200_FASTWALK(3) VB200_FASTWALK(5)
#undef VB200_FASTWALK
    ctx->launches++;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) *rc = fail(ctx, VB200_ERR_CUDA, "fast bin walk failed to launch: %s", cudaGetErrorString(e));
    return true;
}

// ---- weighted Russian roulette (rr_integral_region / rr_error_region): host side of the kernels above --------------------------
int region_total_errors(vb200_ctx* ctx, const vb200_regions* r, const BinWalk& w, float* rerr) {
    if (r->SL <= 0) return fail(ctx, VB200_ERR_UNSUPPORTED, "rr_error_region needs a nested rule (Region::error(), region.h:420)");
    if (r->sd > 729) return fail(ctx, VB200_ERR_UNSUPPORTED, "rr_error_region: %d samples per region exceed the kernel's scratch", r->sd);
    const unsigned grid = unsigned((r->count + 63) / 64);
    if (r->SH == 3 && r->SL == 2) region_total_error_kernel<3, 2><<<grid, 64, 0, ctx->stream>>>(r->count, r->capacity, r->sd, r->data, w.volume, rerr);
    else if (r->SH == 5 && r->SL == 3) region_total_error_kernel<5, 3><<<grid, 64, 0, ctx->stream>>>(r->count, r->capacity, r->sd, r->data, w.volume, rerr);
    else return fail(ctx, VB200_ERR_UNSUPPORTED, "rr_error_region: no kernel for the nested pair (%d,%d)", r->SH, r->SL);
    ctx->launches++;
    VB200_CUDA(ctx, cudaGetLastError());
    return VB200_OK;
}

// pdf patches [3^db][cap] for rr_pdf_region: the region samples folded over the non-binned dimensions D-1 .. db with the pdf line
// integral over [0,1] (the same for every bin, like BinWalk::patches).  The caller frees *out with dfree.
int walk_pdf_patches(vb200_ctx* ctx, const vb200_regions* r, const BinWalk& w, float** out) {
    *out = nullptr;
    if (r->SH != 3) return fail(ctx, VB200_ERR_UNSUPPORTED, "rr_pdf_region needs a Simpson-based rule (only Simpson defines pdf_integral_subrange, rules.h:151)");
    const int D = r->dim, db = w.db; const uint64_t n = r->count, cap = r->capacity;
    float* bufs[2] = {nullptr, nullptr};
    const float* cur = r->data;
    uint64_t biggest = 1; for (int i = 0; i < D - 1; ++i) biggest *= 3;
    int which = 0;
    auto cleanup = [&] (float* keep) { for (float* b : bufs) if (b && b != keep) dfree(ctx, b); };
    for (int m = D; m > db; --m) {
        int lower = 1; for (int i = 0; i < m - 1; ++i) lower *= 3;
        if (!bufs[which] && dmalloc(ctx, &bufs[which], biggest * cap * sizeof(float)) != cudaSuccess) { cudaGetLastError(); cleanup(nullptr); return fail(ctx,
                VB200_ERR_NOMEM, "cudaMalloc failed (pdf patches)"); }
        dim3 grid(unsigned((n + 127) / 128), unsigned(lower));
        pdf_fold_last_dim_kernel<float><<<grid, 128, 0, ctx->stream>>>(n, cap, lower, cur, bufs[which]);
        ctx->launches++;
        cur = bufs[which]; which ^= 1;
    }
    if (cur == r->data) {      // every dimension is binned: the pdf patch is the sample table itself
        if (dmalloc(ctx, &bufs[0], uint64_t(r->sd) * cap * sizeof(float)) != cudaSuccess) { cudaGetLastError(); return fail(ctx, VB200_ERR_NOMEM, "cudaMalloc failed (pdf patches)"); }
        cudaMemcpyAsync(bufs[0], r->data, uint64_t(r->sd) * cap * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream);
        cur = bufs[0];
    }
    cleanup(const_cast<float*>(cur));
    *out = const_cast<float*>(cur);
    VB200_CUDA(ctx, cudaGetLastError());
    return VB200_OK;
}

static double rr_floor_factor(int policy) { return policy == VB200_RR_PDF ? double(0.01f) : 0.01; }    // rr_pdf_region keeps factor_prob as a float (:111)

int walk_rr_pass(vb200_ctx* ctx, const vb200_regions* r, const BinWalk& w, const vb200_domain& domain, uint64_t begin, uint64_t end, uint64_t base, int policy, int pass,
                 const float* rerr, const float* pdf_patches, const uint32_t* count, double* wsum, double* csum, uint32_t spp, const uint32_t* raw, uint32_t* chosen) {
    const DomT<float> dom = to_dom(domain);
    const TileGeom g = make_geom<float>(w, dom);
    if (policy == VB200_RR_PDF && (w.S != 3 || !pdf_patches)) return fail(ctx, VB200_ERR_UNSUPPORTED,
            "rr_pdf_region needs a Simpson-based rule (only Simpson defines pdf_integral_subrange, rules.h:151)");
    const float* patches = policy == VB200_RR_PDF ? pdf_patches : w.patches;
    const double ff = rr_floor_factor(policy);
#define VB200_RRW(SS, DD) if (w.S == SS && w.db == DD) { walk_rr_kernel<SS, DD><<<unsigned(w.ntiles), 256, 0, ctx->stream>>>(g, dom, w.cap, begin, end, base, r->dim, policy, pass, ff, \
        patches, r->rmin, r->rmax, w.volume, w.pstart, w.pend, w.tile_offset, w.tile_list, rerr, count, wsum, csum, spp, raw, chosen); ctx->launches++; VB200_CUDA(ctx, cudaGetLastError()); return VB200_OK; }
    VB200_RRW(2, 1) VB200_RRW(2, 2) VB200_RRW(2, 3) VB200_RRW(3, 1) VB200_RRW(3, 2) VB200_RRW(3, 3) VB200_RRW(5, 1) VB200_RRW(5, 2) VB200_RRW(5, 3)
#undef VB200_RRW
    return fail(ctx, VB200_ERR_UNSUPPORTED, "no weighted roulette walk for rule with %d samples and %d binned dimensions", w.S, w.db);
}

int rr_factors(vb200_ctx* ctx, const vb200_regions* r, const BinWalk& w, const vb200_domain& domain, uint64_t s0, uint64_t nb, uint64_t base, int policy, uint32_t spp,
               const float* rerr, const float* pdf_patches, const uint32_t* count, const double* wsum, const double* csum, const uint32_t* chosen, double* rrf) {
    const DomT<float> dom = to_dom(domain);
    if (policy == VB200_RR_PDF && (w.S != 3 || !pdf_patches)) return fail(ctx, VB200_ERR_UNSUPPORTED, "rr_pdf_region needs a Simpson-based rule");
    const float* patches = policy == VB200_RR_PDF ? pdf_patches : w.patches;
    const double ff = rr_floor_factor(policy);
    const unsigned grid = unsigned((nb * spp + 127) / 128);
#define VB200_RRF(SS, DD) if (w.S == SS && w.db == DD) { rr_factor_kernel<SS, DD><<<grid, 128, 0, ctx->stream>>>(dom, w.cap, s0, nb, base, r->dim, policy, ff, spp, \
        patches, r->rmin, r->rmax, w.volume, rerr, count, wsum, csum, chosen, rrf); ctx->launches++; VB200_CUDA(ctx, cudaGetLastError()); return VB200_OK; }
    VB200_RRF(2, 1) VB200_RRF(2, 2) VB200_RRF(2, 3) VB200_RRF(3, 1) VB200_RRF(3, 2) VB200_RRF(3, 3) VB200_RRF(5, 1) VB200_RRF(5, 2) VB200_RRF(5, 3)
#undef VB200_RRF
    return fail(ctx, VB200_ERR_UNSUPPORTED, "no weighted roulette kernel for rule with %d samples and %d binned dimensions", w.S, w.db);
}

// the two scalar types the library computes in
template int walk_build_t<float>(vb200_ctx*, const vb200_regions*, const DomT<float>&, uint64_t, uint64_t, BinWalkT<float>*);
template int walk_build_t<double>(vb200_ctx*, const vb200_regions*, const DomT<double>&, uint64_t, uint64_t, BinWalkT<double>*);
template void walk_free_t<float>(BinWalkT<float>*);
template void walk_free_t<double>(BinWalkT<double>*);
template int walk_accumulate_t<float>(vb200_ctx*, const vb200_regions*, const BinWalkT<float>&, const DomT<float>&, uint64_t, uint64_t, int, float*, float*, uint32_t*);
template int walk_accumulate_t<double>(vb200_ctx*, const vb200_regions*, const BinWalkT<double>&, const DomT<double>&, uint64_t, uint64_t, int, double*, double*, uint32_t*);

} // namespace vb200

// Range<double,DIM>: RegionsIntegratorSequential with Float = value_type = double — every fold, the bin boxes and the running
// '+=' are evaluated in double, as upstream
extern "C" int vb200_regions_integrate_bins_f64(vb200_ctx* ctx, const vb200_regions* r, const vb200_domain_f64* domain, const vb200_shard* shard,
                                                double* bins, int bins_mem) {
    if (!ctx || !r || !domain || !bins) return fail(ctx, VB200_ERR_INVALID, "NULL argument");
    VB200_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!r->f64) return fail(ctx, VB200_ERR_INVALID, "region table holds float values: use vb200_regions_integrate_bins");
    if (domain->dim != r->dim || domain->dimbins < 1 || domain->dimbins > VB200_MAX_DIMBINS || domain->dimbins > domain->dim)
        return fail(ctx, VB200_ERR_INVALID, "range has %d dimensions (%d binned), regions have %d", domain->dim, domain->dimbins, r->dim);
    for (int i = 0; i < domain->dimbins; ++i) if (domain->res[i] == 0 || domain->res[i] > 0xffffffffull) return fail(ctx, VB200_ERR_INVALID, "resolution[%d] invalid", i);
    const DomT<double> dom = to_dom(*domain);
    const uint64_t total = nbins_of(dom);
    uint64_t begin, end; vb200_shard s = shard ? *shard : vb200_shard{0, 0};
    int rc = resolve_shard(ctx, s, total, &begin, &end); if (rc) return rc;
    if (begin == end || r->count == 0) return VB200_OK;
    const uint64_t n = end - begin;
    double* dev_base = bins;
    if (bins_mem == VB200_HOST) {
        void* d = nullptr; rc = reserve(ctx, 0, n * sizeof(double), &d); if (rc) return rc;
        VB200_CUDA(ctx, cudaMemcpyAsync(d, bins + begin, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));     // '+=' continues from the current contents
        dev_base = static_cast<double*>(d) - begin;
    } else if (bins_mem != VB200_DEVICE) return fail(ctx, VB200_ERR_INVALID, "bad memory-space flag %d", bins_mem);
    BinWalkT<double> w;
    rc = walk_build_t<double>(ctx, r, dom, begin, end, &w); if (rc) return rc;
    rc = walk_accumulate_t<double>(ctx, r, w, dom, begin, end, 0, dev_base, nullptr, nullptr);
    if (!rc && bins_mem == VB200_HOST) {
        cudaError_t e = cudaMemcpyAsync(bins + begin, dev_base + begin, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) rc = fail(ctx, VB200_ERR_CUDA, "copy back failed: %s", cudaGetErrorString(e));
    }
    if (!rc) { cudaError_t e = cudaStreamSynchronize(ctx->stream); if (e != cudaSuccess) rc = fail(ctx, VB200_ERR_CUDA, "region->bin integration (f64) failed: %s", cudaGetErrorString(e)); }
    walk_free_t<double>(&w);
    return rc;
}

// ---- Steps<Q,N> composite rules: one region over the whole range, one thread per bin ---------------------------------------------
namespace {
// Steps::subrange (rules.h:359-384) over a line whose elements are produced on demand (the folds of the dimensions above it)
template<int Q, class Elem>
__device__ __forceinline__ float steps_subrange(float a, float b, uint32_t N, Elem&& elem) {
    const float fN = float(N);
    const float aN = R::fm(a, fN), bN = R::fm(b, fN);
    uint32_t ia = uint32_t(aN > 0.0f ? (unsigned long long)aN : 0ull), ib = uint32_t(bN > 0.0f ? (unsigned long long)bN : 0ull);
    if (ia > N - 1) ia = N - 1;
    if (ib > N - 1) ib = N - 1;
    const float a_local = R::fs(aN, float(ia)), b_local = R::fs(bN, float(ib));
    float l[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) l[k] = elem(ia * uint32_t(Q - 1) + uint32_t(k));
    if (ia == ib) return R::fd(R::subrange<Q, float>(a_local, b_local, l), fN);
    float sol = R::fd(R::subrange<Q, float>(a_local, 1.0f, l), fN);
    for (uint32_t i = ia + 1; i < ib; ++i) {
#pragma unroll
        for (int k = 0; k < Q; ++k) l[k] = elem(i * uint32_t(Q - 1) + uint32_t(k));
        sol = R::fa(sol, R::fd(R::apply<Q, float>(l), fN));
    }
#pragma unroll
    for (int k = 0; k < Q; ++k) l[k] = elem(ib * uint32_t(Q - 1) + uint32_t(k));
    return R::fa(sol, R::fd(R::subrange<Q, float>(0.0f, b_local, l), fN));
}
// Region::sub_last (region.h:141-151): subrange folded over dimension D-1 first ... dimension 0 last; evaluated depth first
template<int Q, int D, int LEVEL> struct StepsFold {
    struct Elem {
        const float* data; uint64_t base, stride; uint32_t S, N; const float* a; const float* b;
        __device__ __forceinline__ float operator()(uint32_t j) const { return StepsFold<Q, D, LEVEL + 1>::eval(data, base + uint64_t(j) * stride, S, N, a, b); }
    };
    // NOT inlined: every level calls the next one from 3*Q sites, so inlining the recursion would multiply the code by (3Q)^D
    __device__ __noinline__ static float eval(const float* __restrict__ data, uint64_t base, uint32_t S, uint32_t N, const float* a, const float* b) {
        uint64_t stride = 1;
#pragma unroll
        for (int i = 0; i < LEVEL; ++i) stride *= S;
        return steps_subrange<Q>(a[LEVEL], b[LEVEL], N, Elem{data, base, stride, S, N, a, b});
    }
};
template<int Q, int D> struct StepsFold<Q, D, D> {
    __device__ __forceinline__ static float eval(const float* __restrict__ data, uint64_t base, uint32_t, uint32_t, const float*, const float*) { return __ldg(data + base); }
};

template<int Q, int D>
__global__ void __launch_bounds__(128) steps_integrate_kernel(vb200_domain dom, uint64_t begin, uint64_t end, uint64_t nbins_total, uint32_t S, uint32_t N,
                                                              const float* __restrict__ rmin, const float* __restrict__ rmax, const float* __restrict__ data, float* __restrict__ out) {
    const uint64_t bin = begin + uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (bin >= end) return;
    uint32_t pos[VB200_MAX_DIMBINS]; { uint64_t q = bin; for (int d = 0; d < dom.dimbins; ++d) { pos[d] = uint32_t(q % dom.res[d]); q /= dom.res[d]; } }
    float a[D], b[D]; float volume = 1.0f; bool empty = false;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const float lo = rmin[d], hi = rmax[d];                                  // the region = the whole range (capacity 1)
        volume = R::fm(volume, R::fs(hi, lo));                                     // Range::volume (range.h:21-25)
        float ba = dom.rmin[d], bb = dom.rmax[d];
        if (d < dom.dimbins) {
            // pixels_in_region (region.h:454-463): bins outside [start,end) are never visited
            const float res = float(dom.res[d]), ext = R::fs(dom.rmax[d], dom.rmin[d]);
            const float fs_ = R::fd(R::fm(res, R::fs(lo, dom.rmin[d])), ext), fe_ = R::fa(0.99f, R::fd(R::fm(res, R::fs(hi, dom.rmin[d])), ext));
            uint64_t s = fs_ > 0.0f ? uint64_t(fs_) : 0ull, e = fe_ > 0.0f ? uint64_t(fe_) : 0ull;
            if (e > dom.res[d]) e = dom.res[d];
            if (e < s + 1) e = s + 1;
            if (pos[d] < s || pos[d] >= e) empty = true;
            ba = R::fa(dom.rmin[d], R::fm(float(pos[d]), dom.drange[d])); bb = R::fa(dom.rmin[d], R::fm(float(pos[d] + 1u), dom.drange[d]));
        }
        const float ia = fmaxf(ba, lo), ib = fmaxf(ia, fminf(bb, hi));             // Range::intersection (range.h:92-101)
        if (ia >= ib) empty = true;
        a[d] = R::pos_in_range(lo, hi, ia); b[d] = R::pos_in_range(lo, hi, ib);
    }
    if (empty) return;
    const float integral = R::fm(volume, StepsFold<Q, D, 0>::eval(data, 0, S, N, a, b));
    out[bin] = R::d2f(R::da(double(out[bin]), R::dm(double(nbins_total), double(integral))));      // regions-integrator-sequential.h:54
}

template<int Q>
int steps_dispatch(vb200_ctx* ctx, const vb200_regions* r, const vb200_domain& dom, uint64_t begin, uint64_t end, uint64_t total, float* out) {
    const unsigned grid = unsigned((end - begin + 127) / 128);
    const uint32_t S = uint32_t(r->SH), N = uint32_t(r->rule & 0xffff);
    switch (r->dim) {
        case 1: steps_integrate_kernel<Q, 1><<<grid, 128, 0, ctx->stream>>>(dom, begin, end, total, S, N, r->rmin, r->rmax, r->data, out); break;
        case 2: steps_integrate_kernel<Q, 2><<<grid, 128, 0, ctx->stream>>>(dom, begin, end, total, S, N, r->rmin, r->rmax, r->data, out); break;
        case 3: steps_integrate_kernel<Q, 3><<<grid, 128, 0, ctx->stream>>>(dom, begin, end, total, S, N, r->rmin, r->rmax, r->data, out); break;
        case 4: steps_integrate_kernel<Q, 4><<<grid, 128, 0, ctx->stream>>>(dom, begin, end, total, S, N, r->rmin, r->rmax, r->data, out); break;
        default: return fail(ctx, VB200_ERR_UNSUPPORTED, "composite (steps) rules are instantiated for 1..4 dimensions");
    }
    ctx->launches++;
    VB200_CUDA(ctx, cudaGetLastError());
    return VB200_OK;
}
int steps_integrate(vb200_ctx* ctx, const vb200_regions* r, const vb200_domain& dom, uint64_t begin, uint64_t end, uint64_t total, float* out) {
    if (r->count != 1 || r->capacity != 1) return fail(ctx, VB200_ERR_UNSUPPORTED, "composite (steps) rules describe a single region");
    switch ((r->rule >> 16) & 0xff) {
        case 2: return steps_dispatch<2>(ctx, r, dom, begin, end, total, out);
        case 3: return steps_dispatch<3>(ctx, r, dom, begin, end, total, out);
        case 5: return steps_dispatch<5>(ctx, r, dom, begin, end, total, out);
    }
    return fail(ctx, VB200_ERR_INVALID, "bad steps rule %d", r->rule);
}
}

extern "C" int vb200_regions_integrate_bins(vb200_ctx* ctx, const vb200_regions* r, const vb200_domain* domain, const vb200_shard* shard,
                                            float* bins, int bins_mem) {
    if (!ctx || !r || !domain || !bins) return fail(ctx, VB200_ERR_INVALID, "NULL argument");
    VB200_CUDA(ctx, cudaSetDevice(ctx->device));
    if (r->f64) return fail(ctx, VB200_ERR_INVALID, "region table holds double values: use vb200_regions_integrate_bins_f64");
    int rc = check_domain(ctx, *domain, r->dim); if (rc) return rc;
    const vb200_domain dom = finish_domain(*domain);
    const uint64_t total = nbins_of(dom);
    uint64_t begin, end; vb200_shard s = shard ? *shard : vb200_shard{0, 0};
    rc = resolve_shard(ctx, s, total, &begin, &end); if (rc) return rc;
    if (begin == end || r->count == 0) return VB200_OK;
    // '+=' continues from the bins' current contents in the reference's float(double(acc)+...) chain: upload them
    BinStage st; rc = stage_bins_in(ctx, bins, bins_mem, begin, end, /*upload=*/true, &st); if (rc) return rc;
    if (VB200_RULE_IS_STEPS(r->rule)) {
        rc = steps_integrate(ctx, r, dom, begin, end, total, st.dev_base);
        if (!rc) rc = stage_bins_out(ctx, st);
        if (!rc && !st.staged) { cudaError_t e = cudaStreamSynchronize(ctx->stream); if (e != cudaSuccess) rc = fail(ctx, VB200_ERR_CUDA, "steps integration failed: %s", cudaGetErrorString(e)); }
        return rc;
    }
    BinWalk w;
    rc = walk_build(ctx, r, dom, begin, end, &w); if (rc) return rc;
    rc = walk_accumulate(ctx, r, w, dom, begin, end, 0, st.dev_base, nullptr, nullptr);
    if (!rc) rc = stage_bins_out(ctx, st);
    if (!rc && !st.staged) { cudaError_t e = cudaStreamSynchronize(ctx->stream); if (e != cudaSuccess) rc = fail(ctx, VB200_ERR_CUDA, "region->bin integration failed: %s", cudaGetErrorString(e)); }
    walk_free(&w);
    return rc;
}

// ---- adaptive refinement ------------------------------------------------------------------------------------------------
// heap-array order -> region table (SoA): region h of the output is the region the h-th heap entry points at
// (the reference returns the heap vector as is, regions-generator-adaptive-heap.h:44).  T = scalar type of the table; KEY64: 16-byte
// heap entries with double keys (double tables and error_heuristic_mixed), whose key goes to err[] rounded to T.
template<class T, bool KEY64>
__global__ void compact_heap_order_kernel(uint64_t n, uint64_t cap_out, int dim, int sd, const void* __restrict__ heap_,
                                          const T* __restrict__ range, const T* __restrict__ data,
                                          T* __restrict__ rmin, T* __restrict__ rmax, T* __restrict__ odata,
                                          T* __restrict__ err, uint32_t* __restrict__ errdim) {
    const uint64_t h = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (h >= n) return;
    uint64_t id; uint32_t ed; T key;
    if constexpr (KEY64) {
        const ulonglong2 e = static_cast<const ulonglong2*>(heap_)[h];
        id = e.y & 0x0fffffffull; ed = uint32_t(e.y >> 28) & 0xfu; key = T(__longlong_as_double(static_cast<long long>(e.x)));
    } else {
        const unsigned long long e = static_cast<const unsigned long long*>(heap_)[h];
        id = (e >> 32) & 0x0fffffffull; ed = uint32_t(e >> 60); key = T(__uint_as_float(unsigned(e)));
    }
    if (blockIdx.y == 0) {
        err[h] = key; errdim[h] = ed;
        for (int d = 0; d < dim; ++d) { rmin[uint64_t(d) * cap_out + h] = range[id * uint64_t(2 * dim) + d]; rmax[uint64_t(d) * cap_out + h] = range[id * uint64_t(2 * dim) + dim + d]; }
    }
    for (int k = blockIdx.y; k < sd; k += gridDim.y) odata[uint64_t(k) * cap_out + h] = data[id * uint64_t(sd) + k];
}

// exact greedy mode (batch = 1) for float and double tables; `dmin/dmax` are the range in the table's type
template<class T>
static int generate_greedy_t(vb200_ctx* ctx, const vb200_integrand* f, int rule, int heuristic, int metric, double size_weight, uint64_t iterations,
                             const vb200_mixed_heuristic& mixed, const T* dmin, const T* dmax, vb200_regions** out) {
    constexpr bool F64 = sizeof(T) == 8;
    const bool key64 = F64 || heuristic == VB200_HEURISTIC_MIXED;
    vb200_regions* r = nullptr;
    const uint64_t n = iterations + 1, cap = 2 * iterations + 1;
    int rc = regions_alloc(ctx, f->dim, rule, n, &r, F64); if (rc) return rc;
    const int D = f->dim; const uint64_t sd = uint64_t(r->sd);
    T *range = nullptr, *data = nullptr, *err = nullptr; double* key = nullptr; void* heap = nullptr; uint64_t* hsize = nullptr;
    auto cleanup = [&] () { dfree(ctx, range); dfree(ctx, data); dfree(ctx, err); dfree(ctx, key); dfree(ctx, heap); dfree(ctx, hsize); };
    auto bail = [&] (int code) { cleanup(); vb200_regions_free(r); return code; };
    if (dmalloc(ctx, &range, cap * 2 * D * sizeof(T)) != cudaSuccess || dmalloc(ctx, &data, cap * sd * sizeof(T)) != cudaSuccess ||
        dmalloc(ctx, &err, cap * sizeof(T)) != cudaSuccess || (key64 && dmalloc(ctx, &key, cap * sizeof(double)) != cudaSuccess) ||
        dmalloc_bytes(ctx, &heap, (n + 1) * (key64 ? 16 : 8)) != cudaSuccess ||
        dmalloc(ctx, &hsize, sizeof(uint64_t)) != cudaSuccess) { cudaGetLastError(); return bail(fail(ctx, VB200_ERR_NOMEM,
                "working set of the greedy refinement (%llu region slots) does not fit", (unsigned long long)cap)); }
    vb200_greedy_launch a; std::memset(&a, 0, sizeof(a));
    a.dim = D; a.rule = rule; a.heuristic = heuristic; a.metric = metric; a.size_weight = size_weight;
    a.iterations = iterations; a.capacity = cap; a.range = range; a.data = data; a.err = err; a.heap = heap; a.heap_size = hsize; a.key64 = key;
    a.f64 = F64 ? 1 : 0; a.metric_rest = mixed.metric_rest; a.mixed_dimension = mixed.dimension; a.mixed_bins_weight = mixed.bins_weight;
    a.mixed_threshold_bins = mixed.size_threshold_bins; a.mixed_threshold_rest = mixed.size_threshold_rest; a.mixed_error_increase = mixed.error_increase_factor;
    for (int d = 0; d < D; ++d) {
        if (F64) { a.range_min64[d] = double(dmin[d]); a.range_max64[d] = double(dmax[d]); }
        else { a.range_min[d] = float(dmin[d]); a.range_max[d] = float(dmax[d]); }
    }
    rc = call_thunk(ctx, f, VB200_K_ADAPTIVE_EXACT, &a); if (rc) return bail(rc);
    uint64_t got = 0;
    if (cudaMemcpyAsync(&got, hsize, sizeof(got), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        return bail(fail(ctx, VB200_ERR_CUDA, "greedy refinement kernel failed: %s", cudaGetErrorString(cudaGetLastError())));
    if (got != n) return bail(fail(ctx, VB200_ERR_CUDA, "greedy refinement ended with %llu regions, expected %llu", (unsigned long long)got, (unsigned long long)n));
    dim3 grid(unsigned((n + 255) / 256), unsigned(sd < 32 ? sd : 32));
    if (key64) compact_heap_order_kernel<T, true><<<grid, 256, 0, ctx->stream>>>(n, n, D, int(sd), heap, range, data, RegCols<T>::rmin(r), RegCols<T>::rmax(r),
            RegCols<T>::data(r), RegCols<T>::err(r), r->errdim);
    else compact_heap_order_kernel<T, false><<<grid, 256, 0, ctx->stream>>>(n, n, D, int(sd), heap, range, data, RegCols<T>::rmin(r), RegCols<T>::rmax(r),
            RegCols<T>::data(r), RegCols<T>::err(r), r->errdim);
    ctx->launches++;
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        return bail(fail(ctx, VB200_ERR_CUDA, "heap-order compaction failed: %s", cudaGetErrorString(cudaGetLastError())));
    cleanup();
    r->count = n;
    *out = r;
    return VB200_OK;
}
static int check_mixed(vb200_ctx* ctx, int heuristic, const vb200_mixed_heuristic& m) {
    if (heuristic != VB200_HEURISTIC_MIXED) return VB200_OK;
    if (m.metric_rest != VB200_METRIC_ABSOLUTE && m.metric_rest != VB200_METRIC_RELATIVE) return fail(ctx, VB200_ERR_INVALID, "error_heuristic_mixed: unknown rest metric %d", m.metric_rest);
    if (m.dimension < 0) return fail(ctx, VB200_ERR_INVALID, "error_heuristic_mixed: dimension %d", m.dimension);
    return VB200_OK;
}

extern "C" int vb200_regions_generate_tolerance(vb200_ctx* ctx, const vb200_integrand* f, const vb200_tolerance_params* p, vb200_regions** out) {
    if (!ctx || !f || !p || !out) return fail(ctx, VB200_ERR_INVALID, "NULL argument");
    VB200_CUDA(ctx, cudaSetDevice(ctx->device));
    if (f->dim <= 0 || p->domain.dim != f->dim) return fail(ctx, VB200_ERR_INVALID, "range has %d dimensions, integrand takes %d", p->domain.dim, f->dim);
    int SH, SL;
    if (rule_samples(p->rule, &SH, &SL) || SL == 0) return fail(ctx, VB200_ERR_INVALID, "adaptive refinement needs a nested(high,low) rule (got %d)", p->rule);
    if (p->heuristic != VB200_HEURISTIC_DEFAULT && p->heuristic != VB200_HEURISTIC_SIZE) return fail(ctx, VB200_ERR_INVALID, "unknown heuristic %d", p->heuristic);
    if (p->metric != VB200_METRIC_ABSOLUTE && p->metric != VB200_METRIC_RELATIVE) return fail(ctx, VB200_ERR_INVALID, "unknown metric %d", p->metric);
    if (!(p->tolerance > 0.0f)) return fail(ctx, VB200_ERR_INVALID, "tolerance %g must be positive (no region's error is below it otherwise)", double(p->tolerance));
    return generate_tolerance(ctx, f, p, out);
}

extern "C" int vb200_regions_generate_adaptive(vb200_ctx* ctx, const vb200_integrand* f, const vb200_adaptive_params* p, vb200_regions** out) {
    if (!ctx || !f || !p || !out) return fail(ctx, VB200_ERR_INVALID, "NULL argument");
    VB200_CUDA(ctx, cudaSetDevice(ctx->device));
    if (f->dim <= 0 || p->domain.dim != f->dim) return fail(ctx, VB200_ERR_INVALID, "range has %d dimensions, integrand takes %d", p->domain.dim, f->dim);
    int SH, SL;
    if (rule_samples(p->rule, &SH, &SL) || SL == 0) return fail(ctx, VB200_ERR_INVALID, "adaptive refinement needs a nested(high,low) rule (got %d)", p->rule);
    if (p->heuristic != VB200_HEURISTIC_DEFAULT && p->heuristic != VB200_HEURISTIC_SIZE && p->heuristic != VB200_HEURISTIC_MIXED) return fail(ctx, VB200_ERR_INVALID,
            "unknown heuristic %d", p->heuristic);
    if (p->metric != VB200_METRIC_ABSOLUTE && p->metric != VB200_METRIC_RELATIVE) return fail(ctx, VB200_ERR_INVALID, "unknown metric %d", p->metric);
    if (int rcm = check_mixed(ctx, p->heuristic, p->mixed)) return rcm;
    if (f->flags & VB200_INTEGRAND_F64) return fail(ctx, VB200_ERR_INVALID, "double-precision integrand: use vb200_regions_generate_adaptive_f64");
    if (p->iterations >= (1ull << 27)) return fail(ctx, VB200_ERR_UNSUPPORTED, "more than 2^27 iterations");
    if (p->batch == 1) return generate_greedy_t<float>(ctx, f, p->rule, p->heuristic, p->metric, p->size_weight, p->iterations, p->mixed, p->domain.rmin, p->domain.rmax, out);
    if (p->batch < 0) return fail(ctx, VB200_ERR_INVALID, "batch=%d", p->batch);
    return generate_batched(ctx, f, p, out);
}

extern "C" int vb200_regions_generate_adaptive_f64(vb200_ctx* ctx, const vb200_integrand* f, const vb200_adaptive_params_f64* p, vb200_regions** out) {
    if (!ctx || !f || !p || !out) return fail(ctx, VB200_ERR_INVALID, "NULL argument");
    VB200_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!(f->flags & VB200_INTEGRAND_F64)) return fail(ctx, VB200_ERR_INVALID, "vb200_regions_generate_adaptive_f64 needs a double-precision integrand");
    if (f->dim <= 0 || p->domain.dim != f->dim) return fail(ctx, VB200_ERR_INVALID, "range has %d dimensions, integrand takes %d", p->domain.dim, f->dim);
    int SH, SL;
    if (rule_samples(p->rule, &SH, &SL) || SL == 0) return fail(ctx, VB200_ERR_INVALID, "adaptive refinement needs a nested(high,low) rule (got %d)", p->rule);
    if (p->heuristic != VB200_HEURISTIC_DEFAULT && p->heuristic != VB200_HEURISTIC_SIZE && p->heuristic != VB200_HEURISTIC_MIXED) return fail(ctx, VB200_ERR_INVALID,
            "unknown heuristic %d", p->heuristic);
    if (p->metric != VB200_METRIC_ABSOLUTE && p->metric != VB200_METRIC_RELATIVE) return fail(ctx, VB200_ERR_INVALID, "unknown metric %d", p->metric);
    if (int rcm = check_mixed(ctx, p->heuristic, p->mixed)) return rcm;
    if (p->iterations >= (1ull << 27)) return fail(ctx, VB200_ERR_UNSUPPORTED, "more than 2^27 iterations");
    if (p->batch != 1) return fail(ctx, VB200_ERR_UNSUPPORTED, "double-precision adaptive generation runs in the exact greedy mode only (batch = 1)");
    return generate_greedy_t<double>(ctx, f, p->rule, p->heuristic, p->metric, p->size_weight, p->iterations, p->mixed, p->domain.rmin, p->domain.rmax, out);
}
