// libviltrum_b200.so — the multi-GPU side of the C ABI (include/viltrum_b200.h "multi-GPU"): one process per GPU, an optional NCCL
// communicator owned by the context.  SURVEY.md §8(e): every per-bin path shards over independent bins and needs NO collective; the two
// exchanges the design has are
//   * the split-sample mode of vb200_monte_carlo (few bins, many samples: every rank draws its share of the sample counter and the
//     partial grids are summed — ncclAllReduce over NVLink; reference monte-carlo.h:39-63 is the single-process loop it shards), and
//   * the region-table broadcast in front of the control-variate residual pass (BASELINE configs[3]; reference
//     regions-integrator-parallel-variance-reduction.h:53-63 builds its per-bin lists from ONE table), for callers that generate on one
//     rank instead of on all of them.
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy a host process already carries — PyTorch's — or the system's), so the
// library neither links nor requires NCCL for single-GPU use.
#include "context.h"
#include "regions.h"
#include <dlfcn.h>
#include <cstring>
#include <cstdlib>

namespace {

struct NcclUniqueId { char internal[128]; };      // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128), passed by value to ncclCommInitRank
typedef int (*GetUniqueIdFn)(NcclUniqueId*);
typedef int (*CommInitRankFn)(void**, int, NcclUniqueId, int);
typedef int (*CommDestroyFn)(void*);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*BroadcastFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*GroupFn)(void);
typedef const char* (*ErrorStringFn)(int);
typedef int (*GetVersionFn)(int*);
constexpr int kNcclUint8 = 1, kNcclFloat32 = 7, kNcclSum = 0;      // ncclDataType_t / ncclRedOp_t (nccl.h)

struct Nccl {
    void* handle = nullptr;
    GetUniqueIdFn get_unique_id = nullptr; CommInitRankFn comm_init_rank = nullptr; CommDestroyFn comm_destroy = nullptr;
    AllReduceFn all_reduce = nullptr; BroadcastFn broadcast = nullptr; GroupFn group_start = nullptr, group_end = nullptr;
    ErrorStringFn error_string = nullptr; GetVersionFn get_version = nullptr;
    std::string error;
};

Nccl& nccl() {
    static Nccl n;
    static bool tried = false;
    if (tried) return n;
    tried = true;
    const char* env = std::getenv("VB200_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* name : names) {
        if (!name || !*name) continue;
        n.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);      // a copy already mapped under the same soname (PyTorch's bundled NCCL) is reused
        if (n.handle) break;
        n.error = dlerror();
    }
    if (!n.handle) return n;
    auto sym = [&] (const char* s) -> void* { void* p = dlsym(n.handle, s); if (!p) n.error = std::string("libnccl has no symbol ") + s; return p; };
    n.get_unique_id = reinterpret_cast<GetUniqueIdFn>(sym("ncclGetUniqueId"));
    n.comm_init_rank = reinterpret_cast<CommInitRankFn>(sym("ncclCommInitRank"));
    n.comm_destroy = reinterpret_cast<CommDestroyFn>(sym("ncclCommDestroy"));
    n.all_reduce = reinterpret_cast<AllReduceFn>(sym("ncclAllReduce"));
    n.broadcast = reinterpret_cast<BroadcastFn>(sym("ncclBroadcast"));
    n.group_start = reinterpret_cast<GroupFn>(sym("ncclGroupStart"));
    n.group_end = reinterpret_cast<GroupFn>(sym("ncclGroupEnd"));
    n.error_string = reinterpret_cast<ErrorStringFn>(sym("ncclGetErrorString"));
    n.get_version = reinterpret_cast<GetVersionFn>(sym("ncclGetVersion"));
    if (!n.get_unique_id || !n.comm_init_rank || !n.comm_destroy || !n.all_reduce || !n.broadcast || !n.group_start || !n.group_end || !n.error_string) { dlclose(n.handle); n.handle = nullptr; }
    return n;
}

int need_nccl(vb200_ctx* ctx) {
    if (!nccl().handle) return vb200::fail(ctx, VB200_ERR_UNSUPPORTED, "NCCL is not available (dlopen libnccl.so.2: %s); set VB200_NCCL_LIB", nccl().error.c_str());
    return VB200_OK;
}
#define VB200_NCCL(ctx, call) do { int r__ = (call); if (r__ != 0) \
    return vb200::fail(ctx, VB200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, nccl().error_string(r__), __FILE__, __LINE__); } while (0)

} // namespace

namespace vb200 {
void comm_release(vb200_ctx* ctx) {
    if (ctx->comm && nccl().handle) nccl().comm_destroy(ctx->comm);
    ctx->comm = nullptr; ctx->comm_rank = 0; ctx->comm_size = 1;
}
// sum `count` floats over the communicator, in place, on the context's stream (split-sample mode of vb200_monte_carlo)
int comm_allreduce_sum(vb200_ctx* ctx, float* dev, uint64_t count) {
    if (!ctx->comm) return fail(ctx, VB200_ERR_INVALID, "VB200_MC_ALLREDUCE needs a communicator: call vb200_comm_init first");
    VB200_NCCL(ctx, nccl().all_reduce(dev, dev, size_t(count), kNcclFloat32, kNcclSum, ctx->comm, ctx->stream));
    ctx->launches++;
    return VB200_OK;
}
}

extern "C" int vb200_comm_unique_id(vb200_ctx* ctx, void* id) {
    if (!id) return vb200::fail(ctx, VB200_ERR_INVALID, "NULL argument");
    int rc = need_nccl(ctx); if (rc) return rc;
    NcclUniqueId u; std::memset(&u, 0, sizeof(u));
    VB200_NCCL(ctx, nccl().get_unique_id(&u));
    std::memcpy(id, &u, VB200_COMM_ID_BYTES);
    return VB200_OK;
}

extern "C" int vb200_comm_init(vb200_ctx* ctx, const void* id, int rank, int world) {
    if (!ctx || !id) return vb200::fail(ctx, VB200_ERR_INVALID, "NULL argument");
    if (world < 1 || rank < 0 || rank >= world) return vb200::fail(ctx, VB200_ERR_INVALID, "rank %d outside a world of %d", rank, world);
    int rc = need_nccl(ctx); if (rc) return rc;
    if (ctx->comm) return vb200::fail(ctx, VB200_ERR_INVALID, "the context already owns a communicator (vb200_comm_destroy first)");
    VB200_CUDA(ctx, cudaSetDevice(ctx->device));
    NcclUniqueId u; std::memcpy(&u, id, sizeof(u));
    void* comm = nullptr;
    VB200_NCCL(ctx, nccl().comm_init_rank(&comm, world, u, rank));
    ctx->comm = comm; ctx->comm_rank = rank; ctx->comm_size = world;
    return VB200_OK;
}

extern "C" int vb200_comm_destroy(vb200_ctx* ctx) {
    if (!ctx) return VB200_ERR_INVALID;
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    vb200::comm_release(ctx);
    return VB200_OK;
}
extern "C" int vb200_comm_rank(const vb200_ctx* ctx) { return ctx ? ctx->comm_rank : 0; }
extern "C" int vb200_comm_size(const vb200_ctx* ctx) { return ctx && ctx->comm ? ctx->comm_size : 1; }
extern "C" int vb200_nccl_version(void) { int v = 0; if (nccl().handle && nccl().get_version) nccl().get_version(&v); return v; }

// Region-table broadcast: the root's table goes to every rank over NVLink (five ncclBroadcast calls in one group: the SoA columns);
// the other ranks pass *r == NULL and receive a freshly allocated table of the same shape.
extern "C" int vb200_regions_broadcast(vb200_ctx* ctx, vb200_regions** r, int root) {
    if (!ctx || !r) return vb200::fail(ctx, VB200_ERR_INVALID, "NULL argument");
    if (!ctx->comm) return vb200::fail(ctx, VB200_ERR_INVALID, "vb200_regions_broadcast needs a communicator: call vb200_comm_init first");
    if (root < 0 || root >= ctx->comm_size) return vb200::fail(ctx, VB200_ERR_INVALID, "root %d outside a world of %d", root, ctx->comm_size);
    VB200_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool is_root = ctx->comm_rank == root;
    if (is_root && (!*r || (*r)->f64)) return vb200::fail(ctx, VB200_ERR_INVALID, "the root needs a single-precision table to send");
    if (!is_root && *r) return vb200::fail(ctx, VB200_ERR_INVALID, "non-root ranks pass *r == NULL and receive a new table");
    // shape first: (dim, rule, count) through a 32-byte device buffer
    uint64_t* d_shape = nullptr;
    VB200_CUDA(ctx, vb200::dmalloc(ctx, &d_shape, 4 * sizeof(uint64_t)));
    uint64_t shape[4] = {0, 0, 0, 0};
    if (is_root) { shape[0] = uint64_t((*r)->dim); shape[1] = uint64_t(uint32_t((*r)->rule)); shape[2] = (*r)->count; shape[3] = 0x76623230ull;
                   cudaMemcpyAsync(d_shape, shape, sizeof(shape), cudaMemcpyHostToDevice, ctx->stream); }
    int e = nccl().broadcast(d_shape, d_shape, sizeof(shape), kNcclUint8, root, ctx->comm, ctx->stream);
    cudaError_t ce = cudaMemcpyAsync(shape, d_shape, sizeof(shape), cudaMemcpyDeviceToHost, ctx->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
    vb200::dfree(ctx, d_shape);
    ctx->launches++;
    if (e != 0) return vb200::fail(ctx, VB200_ERR_CUDA, "ncclBroadcast failed: %s", nccl().error_string(e));
    if (ce != cudaSuccess) return vb200::fail(ctx, VB200_ERR_CUDA, "region-table broadcast failed: %s", cudaGetErrorString(ce));
    if (shape[3] != 0x76623230ull || shape[2] == 0) return vb200::fail(ctx, VB200_ERR_INVALID, "region-table broadcast: the root sent no table");
    vb200_regions* t = *r;
    if (!is_root) { int rc = vb200::regions_alloc(ctx, int(shape[0]), int(uint32_t(shape[1])), shape[2], &t); if (rc) return rc; t->count = shape[2]; }
    // the table may have been allocated with spare capacity (generators grow it): the columns are `capacity` apart, so send column by column
    const uint64_t n = t->count, cap = t->capacity;
    e = nccl().group_start();
    for (int d = 0; e == 0 && d < t->dim; ++d) {
        e = nccl().broadcast(t->rmin + uint64_t(d) * cap, t->rmin + uint64_t(d) * cap, n * sizeof(float), kNcclUint8, root, ctx->comm, ctx->stream);
        if (e == 0) e = nccl().broadcast(t->rmax + uint64_t(d) * cap, t->rmax + uint64_t(d) * cap, n * sizeof(float), kNcclUint8, root, ctx->comm, ctx->stream);
    }
    for (int k = 0; e == 0 && k < t->sd; ++k)
        e = nccl().broadcast(t->data + uint64_t(k) * cap, t->data + uint64_t(k) * cap, n * sizeof(float), kNcclUint8, root, ctx->comm, ctx->stream);
    if (e == 0) e = nccl().broadcast(t->err, t->err, n * sizeof(float), kNcclUint8, root, ctx->comm, ctx->stream);
    if (e == 0) e = nccl().broadcast(t->errdim, t->errdim, n * sizeof(uint32_t), kNcclUint8, root, ctx->comm, ctx->stream);
    const int e2 = nccl().group_end();
    ctx->launches++;
    if (e != 0 || e2 != 0) { if (!is_root) vb200_regions_free(t); return vb200::fail(ctx, VB200_ERR_CUDA, "ncclBroadcast failed: %s", nccl().error_string(e != 0 ? e : e2)); }
    *r = t;
    return VB200_OK;
}
