// region tables, region->bin integration and control variates (stubs, being implemented)
#include "context.h"
using namespace vb200;
struct vb200_regions { int dim; };
extern "C" int vb200_regions_generate_adaptive(vb200_ctx* ctx, const vb200_integrand*, const vb200_adaptive_params*, vb200_regions**) { return fail(ctx, VB200_ERR_UNSUPPORTED, "not implemented yet"); }
extern "C" int vb200_regions_generate_single(vb200_ctx* ctx, const vb200_integrand*, const vb200_domain*, int, vb200_regions**) { return fail(ctx, VB200_ERR_UNSUPPORTED, "not implemented yet"); }
extern "C" int vb200_regions_upload(vb200_ctx* ctx, int, int, uint64_t, const float*, const float*, const float*, const uint32_t*, const float*, vb200_regions**) { return fail(ctx, VB200_ERR_UNSUPPORTED, "not implemented yet"); }
extern "C" uint64_t vb200_regions_count(const vb200_regions*) { return 0; }
extern "C" int vb200_regions_dim(const vb200_regions*) { return 0; }
extern "C" int vb200_regions_samples(const vb200_regions*) { return 0; }
extern "C" int vb200_regions_download(vb200_ctx* ctx, const vb200_regions*, float*, float*, float*, uint32_t*, float*) { return fail(ctx, VB200_ERR_UNSUPPORTED, "not implemented yet"); }
extern "C" void vb200_regions_free(vb200_regions*) {}
extern "C" int vb200_regions_integrate_bins(vb200_ctx* ctx, const vb200_regions*, const vb200_domain*, const vb200_shard*, float*, int) { return fail(ctx, VB200_ERR_UNSUPPORTED, "not implemented yet"); }
extern "C" int vb200_cv_integrate(vb200_ctx* ctx, const vb200_integrand*, const vb200_regions*, const vb200_cv_params*, float*, int, uint32_t*, float*) { return fail(ctx, VB200_ERR_UNSUPPORTED, "not implemented yet"); }
extern "C" int vb200_cv_replay(vb200_ctx* ctx, const vb200_integrand*, const vb200_regions*, const vb200_cv_params*, const uint32_t*, const float*, int, float*, int) { return fail(ctx, VB200_ERR_UNSUPPORTED, "not implemented yet"); }
