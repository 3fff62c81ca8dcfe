// BASELINE.json config 2 — per-bin stratified Monte Carlo of a 4-D discontinuous shading integrand, both spellings
// (SURVEY.md §3.2): monte_carlo_per_bin_parallel(spp,seed) ['+='] and integrator_per_bin_parallel(monte_carlo(spp,seed)) ['='].
// The functor is defined HERE, in user code: its kernels are instantiated in this translation unit.
#include <viltrum_b200/viltrum.h>
#include <cstdio>
#include <cstdlib>

template<int K> struct Shade4 {
    __host__ __device__ float operator()(const std::array<float,4>& x) const {
        float a=x[0]-.5f, b=x[1]-.5f;
        float edge=.55f+.35f*(a*a-b*b)+.2f*a*b;
        float vis=(x[2]+.5f*x[3]<edge)?1.0f:0.0f;
        float t=x[2]*(1.0f-x[3]);
        float lobe=1.0f/float(K);
        for (int k=K-2;k>=0;--k) lobe=lobe*t+1.0f/float(k+1);
        float alb=.25f+.75f*x[0]*x[1];
        return vis*lobe*alb;
    }
};
struct Tinted { float gain; __host__ __device__ float operator()(const std::array<float,4>& x) const { return gain*Shade4<8>()(x); } };   // functor with state

int main(int argc, char** argv) {
    const std::size_t w = argc > 1 ? std::atoi(argv[1]) : 256, spp = argc > 2 ? std::atoi(argv[2]) : 64;
    viltrum::tensor<float,2> a({w,w}, 0.0f), b({w,w}, 5.0f), c({w,w}, 0.0f);
    viltrum::integrate(viltrum::monte_carlo_per_bin_parallel(spp, 0), a, a.resolution(), Shade4<64>(), viltrum::range_primary<4>());
    viltrum::integrate(viltrum::integrator_per_bin_parallel(viltrum::monte_carlo(spp, 0)), b, b.resolution(), Shade4<64>(), viltrum::range_primary<4>());
    viltrum::integrate(viltrum::monte_carlo_per_bin_parallel(spp, 1), c, c.resolution(), Tinted{2.0f}, viltrum::range_primary<4>());
    // pinned bins: the kernel accumulates into the tensor's own storage over PCIe (no staging, no host pass); same bits as `a` for the same seed
    viltrum::tensor<float,2> p({w,w}, 0.0f);
    viltrum::b200::pin_bins(p);
    viltrum::integrate(viltrum::monte_carlo_per_bin_parallel(spp, 0), p, p.resolution(), Shade4<64>(), viltrum::range_primary<4>());
    viltrum::b200::unpin_bins(p);
    if (p.raw_data() != a.raw_data()) { std::printf("pinned bins differ from staged bins\n"); return 2; }
    double ma = 0, mb = 0, mc = 0;
    for (float v : a.raw_data()) ma += v; for (float v : b.raw_data()) mb += v; for (float v : c.raw_data()) mc += v;
    ma /= a.size(); mb /= b.size(); mc /= c.size();
    std::printf("mean of bins: %.5f (per_bin_parallel) %.5f (wrapper, overwrote the 5.0 fill) — should be close to 0.14326; tinted %.5f\n", ma, mb, mc);
    if (argc > 3) { FILE* f = std::fopen(argv[3], "wb"); std::fwrite(a.raw_data().data(), 4, a.size(), f); std::fclose(f); }
    return (std::fabs(ma-0.14326) < 2e-3 && std::fabs(mb-0.14326) < 2e-3 && mc > 0.1) ? 0 : 1;
}
