// BASELINE.json config 4 — primary-space piecewise-polynomial control variates + residual Monte Carlo
// (integrator_crespo2021) on a 5-D integrand.
#include <viltrum_b200/viltrum.h>
#include <cstdio>
#include <cstdlib>

template<int K> struct Shade5 {                      // shade4<K>(x0..x3) * (0.5 + x4)
    __host__ __device__ float operator()(const std::array<float,5>& x) const {
        float a=x[0]-.5f, b=x[1]-.5f;
        float edge=.55f+.35f*(a*a-b*b)+.2f*a*b;
        float vis=(x[2]+.5f*x[3]<edge)?1.0f:0.0f;
        float t=x[2]*(1.0f-x[3]);
        float lobe=1.0f/float(K);
        for (int k=K-2;k>=0;--k) lobe=lobe*t+1.0f/float(k+1);
        float alb=.25f+.75f*x[0]*x[1];
        return vis*lobe*alb*(.5f+x[4]);
    }
};

int main(int argc, char** argv) {
    using namespace viltrum;
    const std::size_t w = argc > 1 ? std::atoi(argv[1]) : 64, iterations = argc > 2 ? std::atoi(argv[2]) : 2048, spp = argc > 3 ? std::atoi(argv[3]) : 16;
    tensor<float,2> cv({w,w}, 0.0f), mc({w,w}, 0.0f), ref({w,w}, 0.0f);
    integrate(integrator_crespo2021(iterations, spp, 0), cv, cv.resolution(), Shade5<64>(), range_primary<5>());
    integrate(monte_carlo_per_bin_parallel(spp, 0), mc, mc.resolution(), Shade5<64>(), range_primary<5>());
    integrate(monte_carlo_per_bin_parallel(8192, 1), ref, ref.resolution(), Shade5<64>(), range_primary<5>());
    double m = 0, ecv = 0, emc = 0;
    for (std::size_t i = 0; i < cv.size(); ++i) { m += cv.raw_data()[i]; double r = ref.raw_data()[i]; ecv += (cv.raw_data()[i]-r)*(cv.raw_data()[i]-r); emc += (mc.raw_data()[i]-r)*(mc.raw_data()[i]-r); }
    m /= cv.size();
    std::printf("control variates: mean of bins %.5f should be close to 0.14326; MSE vs 8192-spp reference: CV %.3e, plain MC at the same spp %.3e\n", m, ecv/cv.size(), emc/cv.size());
    // the other Russian-roulette policies, spelled as in the reference's main/compilation-tests/multiple-parameters-2d.cc:62,74
    tensor<float,2> ri({w,w}, 0.0f), re({w,w}, 0.0f);
    integrate(integrator_adaptive_variance_reduction_parallel(nested(simpson,trapezoidal), iterations, rr_integral_region(), cv_optimize_weight(), spp, 0), ri, ri.resolution(), Shade5<64>(), range_primary<5>());
    integrate(integrator_adaptive_variance_reduction_parallel(nested(simpson,trapezoidal), error_heuristic_size(error_metric_relative(),1.e-5), iterations, rr_error_region(), cv_fixed_weight(1.0), region_sampling_uniform(), spp, 0),
              re, re.resolution(), Shade5<64>(), range_primary<5>());
    double mi = 0, me = 0, ei = 0, ee = 0;
    for (std::size_t i = 0; i < ri.size(); ++i) { double r = ref.raw_data()[i]; mi += ri.raw_data()[i]; me += re.raw_data()[i]; ei += (ri.raw_data()[i]-r)*(ri.raw_data()[i]-r); ee += (re.raw_data()[i]-r)*(re.raw_data()[i]-r); }
    mi /= ri.size(); me /= re.size();
    std::printf("rr_integral_region + cv_optimize_weight: mean %.5f, MSE %.3e; rr_error_region + cv_fixed_weight(1): mean %.5f, MSE %.3e\n", mi, ei/ri.size(), me, ee/re.size());
    return (std::fabs(m-0.14326) < 3e-3 && ecv < emc && std::fabs(mi-0.14326) < 3e-3 && std::fabs(me-0.14326) < 3e-3) ? 0 : 1;
}
