// Library-private: device-resident region (leaf) table and the bin-major region walk shared by the Newton-Cotes
// region->bin integrator and the control-variate integrator.
#pragma once
#include "context.h"

// SoA on the device so that a warp touching 32 consecutive regions reads 128-byte lines:
//   rmin/rmax [dim][capacity], data [S^dim][capacity] (sample k of region i at data[k*capacity+i]; k is the reference's
//   multiarray index, dimension 0 fastest: src/multiarray/multiarray.h:20-25), err/errdim [capacity].
struct vb200_regions {
    vb200_ctx* ctx = nullptr;
    int dim = 0, rule = 0, SH = 0, SL = 0, sd = 0;
    uint64_t count = 0, capacity = 0;
    float *rmin = nullptr, *rmax = nullptr, *data = nullptr, *err = nullptr;
    uint32_t* errdim = nullptr;
};

namespace vb200 {

int rule_samples(int rule, int* SH, int* SL);
int regions_alloc(vb200_ctx* ctx, int dim, int rule, uint64_t capacity, vb200_regions** out);
// batched top-k refinement (refine_batched.cu); params already validated
int generate_batched(vb200_ctx* ctx, const vb200_integrand* f, const vb200_adaptive_params* p, vb200_regions** out);

// Per-call acceleration structure for "for every bin, visit the regions that touch it, in table order":
// regions marginalised over the non-binned dimensions (patches), their pixel boxes (region.h:454-463) and per-tile
// ordered region lists.  All device memory, freed by walk_free.
struct BinWalk {
    vb200_ctx* ctx = nullptr;
    int S = 0, db = 0, patch = 0;          // patch = S^db values per region
    uint64_t nregions = 0, cap = 0;
    float* patches = nullptr;              // [patch][cap]
    float* volume = nullptr;               // [cap]  Range::volume of the region, float product in dimension order
    uint32_t* pstart = nullptr;            // [db][cap] pixels_in_region start
    uint32_t* pend = nullptr;              // [db][cap]                  end (exclusive)
    uint32_t tile[3] = {1, 1, 1}, tiles[3] = {1, 1, 1}; uint64_t ntiles = 0;
    uint64_t* tile_offset = nullptr;       // [ntiles+1]
    uint32_t* tile_list = nullptr;         // region ids, ascending inside each tile
    uint64_t pairs = 0;                    // total (tile, region) entries
    float* scratch[2] = {nullptr, nullptr};
};
int walk_build(vb200_ctx* ctx, const vb200_regions* r, const vb200_domain& dom, uint64_t begin, uint64_t end, BinWalk* w);
void walk_free(BinWalk* w);
// mode 0: out[bin] = float(double(out[bin]) + sum over regions)  (RegionsIntegratorSequential '+=', starts from out[bin])
// mode 1: approx[bin-begin] = sum starting from 0, count[bin-begin] = number of regions touching the bin (control variates)
int walk_accumulate(vb200_ctx* ctx, const vb200_regions* r, const BinWalk& w, const vb200_domain& dom, uint64_t begin, uint64_t end,
                    int mode, float* out, float* approx, uint32_t* count);

} // namespace vb200
