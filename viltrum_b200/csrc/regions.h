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
    // double-precision tables (Range<double,DIM>): the same layout in the *64 members, the float members stay null
    bool f64 = false;
    double *rmin64 = nullptr, *rmax64 = nullptr, *data64 = nullptr, *err64 = nullptr;
};
template<class T> struct RegCols;
template<> struct RegCols<float>  { static float*  rmin(const vb200_regions* r) { return r->rmin; }   static float*  rmax(const vb200_regions* r) { return r->rmax; }
                                    static float*  data(const vb200_regions* r) { return r->data; }   static float*  err(const vb200_regions* r) { return r->err; } };
template<> struct RegCols<double> { static double* rmin(const vb200_regions* r) { return r->rmin64; } static double* rmax(const vb200_regions* r) { return r->rmax64; }
                                    static double* data(const vb200_regions* r) { return r->data64; } static double* err(const vb200_regions* r) { return r->err64; } };

namespace vb200 {

int rule_samples(int rule, int* SH, int* SL);
int regions_alloc(vb200_ctx* ctx, int dim, int rule, uint64_t capacity, vb200_regions** out, bool f64 = false);
void orphan_regions(vb200_ctx* ctx);
// batched top-k refinement (refine_batched.cu); params already validated
int generate_batched(vb200_ctx* ctx, const vb200_integrand* f, const vb200_adaptive_params* p, vb200_regions** out);
// tolerance-driven refinement, leaves in the reference's depth-first order (refine_batched.cu); params already validated
int generate_tolerance(vb200_ctx* ctx, const vb200_integrand* f, const vb200_tolerance_params* p, vb200_regions** out);

// Per-call acceleration structure for "for every bin, visit the regions that touch it, in table order":
// regions marginalised over the non-binned dimensions (patches), their pixel boxes (region.h:454-463) and per-tile
// ordered region lists.  All device memory, freed by walk_free.
// integration box + bin grid in the scalar type of the computation (from vb200_domain / vb200_domain_f64)
template<class T> struct DomT {
    int dim = 0, dimbins = 0; T rmin[VB200_MAX_DIM], rmax[VB200_MAX_DIM]; uint64_t res[VB200_MAX_DIMBINS]; T drange[VB200_MAX_DIMBINS];
};
inline DomT<float> to_dom(const vb200_domain& d) {
    DomT<float> o; o.dim = d.dim; o.dimbins = d.dimbins;
    for (int i = 0; i < VB200_MAX_DIM; ++i) { o.rmin[i] = d.rmin[i]; o.rmax[i] = d.rmax[i]; }
    for (int i = 0; i < VB200_MAX_DIMBINS; ++i) { o.res[i] = d.res[i]; o.drange[i] = d.drange[i]; }
    return o;
}
inline DomT<double> to_dom(const vb200_domain_f64& d) {
    DomT<double> o; o.dim = d.dim; o.dimbins = d.dimbins;
    for (int i = 0; i < VB200_MAX_DIM; ++i) { o.rmin[i] = d.rmin[i]; o.rmax[i] = d.rmax[i]; }
    for (int i = 0; i < VB200_MAX_DIMBINS; ++i) { o.res[i] = d.res[i]; o.drange[i] = i < d.dimbins ? (d.rmax[i] - d.rmin[i]) / double(d.res[i]) : 0.0; }
    return o;
}
template<class T> inline uint64_t nbins_of(const DomT<T>& d) { uint64_t n = 1; for (int i = 0; i < d.dimbins; ++i) n *= d.res[i]; return n; }

template<class T> struct BinWalkT {
    vb200_ctx* ctx = nullptr;
    int S = 0, db = 0, patch = 0;          // patch = S^db values per region
    uint64_t nregions = 0, cap = 0;
    T* patches = nullptr;                  // [patch][cap]
    T* volume = nullptr;                   // [cap]  Range::volume of the region, product in T in dimension order
    uint32_t* pstart = nullptr;            // [db][cap] pixels_in_region start
    uint32_t* pend = nullptr;              // [db][cap]                  end (exclusive)
    uint32_t tile[3] = {1, 1, 1}, tiles[3] = {1, 1, 1}; uint64_t ntiles = 0;
    uint64_t* tile_offset = nullptr;       // [ntiles+1]
    uint32_t* tile_list = nullptr;         // region ids, ascending inside each tile
    uint64_t pairs = 0;                    // total (tile, region) entries
    uint64_t max_list = 0;                 // longest tile list
    T* scratch[2] = {nullptr, nullptr};
};
using BinWalk = BinWalkT<float>;
template<class T> int walk_build_t(vb200_ctx* ctx, const vb200_regions* r, const DomT<T>& dom, uint64_t begin, uint64_t end, BinWalkT<T>* w);
template<class T> void walk_free_t(BinWalkT<T>* w);
template<class T> int walk_accumulate_t(vb200_ctx* ctx, const vb200_regions* r, const BinWalkT<T>& w, const DomT<T>& dom, uint64_t begin, uint64_t end,
                                        int mode, T* out, T* approx, uint32_t* count);
inline int walk_build(vb200_ctx* ctx, const vb200_regions* r, const vb200_domain& dom, uint64_t begin, uint64_t end, BinWalk* w) { return walk_build_t<float>(ctx, r, to_dom(dom), begin, end, w); }
inline void walk_free(BinWalk* w) { walk_free_t<float>(w); }
// mode 0: out[bin] = float(double(out[bin]) + sum over regions)  (RegionsIntegratorSequential '+=', starts from out[bin])
// mode 1: approx[bin-begin] = sum starting from 0, count[bin-begin] = number of regions touching the bin (control variates)
inline int walk_accumulate(vb200_ctx* ctx, const vb200_regions* r, const BinWalk& w, const vb200_domain& dom, uint64_t begin, uint64_t end,
                           int mode, float* out, float* approx, uint32_t* count) { return walk_accumulate_t<float>(ctx, r, w, to_dom(dom), begin, end, mode, out, approx, count); }

// throughput form of mode 0/1 in plain fp32 (regions.cu walk_accumulate_fast_kernel): returns false when the shape is not covered
bool walk_accumulate_fast(vb200_ctx* ctx, const vb200_regions* r, const BinWalk& w, const vb200_domain& dom, uint64_t begin, uint64_t end,
                          int mode, float* out, float* approx, uint32_t* count, int* rc);

// Weighted Russian roulette among the regions of a bin (rr_integral_region = policy 1, rr_error_region = policy 2, rr_pdf_region = policy 3;
// reference src/control-variates/region-russian-roulette.h:30-147).  Per-bin arrays are indexed by bin - base.
//   region_total_errors: rerr[r] = Region::error() (policy 2 only)
//   walk_rr_pass: pass 1 -> wsum[bin] = sum of the pair weights; pass 2 -> csum[bin] = sum of the clamped weights; pass 3 -> chosen[j*nb+b]
//                 (sample-major over the slab [begin,end)) = the region whose cumulative clamped weight first exceeds raw[j*nb+b]*2^-32*csum
//   rr_factors: rrf[j*nb+b] = 1.0 / (w'(bin, chosen) / csum[bin])  (1.0 where fewer than two regions touch the bin)
int region_total_errors(vb200_ctx* ctx, const vb200_regions* r, const BinWalk& w, float* rerr);
int walk_pdf_patches(vb200_ctx* ctx, const vb200_regions* r, const BinWalk& w, float** out);   // rr_pdf_region (policy 3): [3^db][cap], freed by the caller (dfree)
int walk_rr_pass(vb200_ctx* ctx, const vb200_regions* r, const BinWalk& w, const vb200_domain& dom, uint64_t begin, uint64_t end, uint64_t base, int policy, int pass,
                 const float* rerr, const float* pdf_patches, const uint32_t* count, double* wsum, double* csum, uint32_t spp, const uint32_t* raw, uint32_t* chosen);
int rr_factors(vb200_ctx* ctx, const vb200_regions* r, const BinWalk& w, const vb200_domain& dom, uint64_t s0, uint64_t nb, uint64_t base, int policy, uint32_t spp,
               const float* rerr, const float* pdf_patches, const uint32_t* count, const double* wsum, const double* csum, const uint32_t* chosen, double* rrf);

} // namespace vb200
