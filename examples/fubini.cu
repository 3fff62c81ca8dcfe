// SURVEY.md §8f rank 2 — the Fubini family, as the reference's own examples spell it:
//   main/doc/integrators.cc:80-100   integrator_fubini<1>(adaptive Newton-Cotes, monte_carlo(32)) on a finite and an infinite integrand,
//                                    integrator_crespo2021_infinite<4>(16, 4, 64)
//   main/compilation-tests/crespo21.cc:31   integrator_crespo2021_infinite<4>(1280, 8, 256) into 10 bins
// User functors (finite: array argument; infinite: generic over a sequence) compiled in this TU.
#include <viltrum_b200/viltrum.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>

struct Slope3 {            // f(x,y,z) = (x+y<1) * (0.5+z): integral over [0,1]^3 = 0.5 ; per-x bins: (1-x_mid) * 1.0
    __host__ __device__ float operator()(const std::array<float,3>& x) const { return ((x[0]+x[1])<1.0f ? 1.0f : 0.0f)*(0.5f+x[2]); }
};
struct Decay {             // reference main/doc/montecarlo-infd.cc:8-22 — geometric series, integral = decay/(1-decay) = 3
    template<typename Seq> __host__ __device__ float operator()(const Seq& seq) const {
        auto x = seq.begin(); float sum = 0.0f, term = 1.0f;
        while ((*x) < 0.75f) { ++x; term *= 2.0f*(*x); ++x; sum += term; }
        return sum;
    }
};
struct Path {              // a bounded random walk: mean 1.013 over the unit square of its first two elements (SURVEY.md App. D "walk")
    template<typename Seq> __host__ __device__ float operator()(const Seq& seq) const {
        auto it = seq.begin(); const float px = *it; ++it; const float py = *it; ++it;
        const float alb = .4f+.5f*(4.0f*px*(1.0f-px))*(.25f+.75f*py);
        float pos = .5f, L = 0.0f;
        while (true) { const float u = *it; ++it; if (u>=alb) break; const float s = *it; ++it; pos = .5f*pos+.5f*s; L += .25f+pos*pos; }
        return L;
    }
};

int main() {
    using namespace viltrum;
    int bad = 0;
    auto adaptive = integrator_adaptive_iterations(nested(simpson, trapezoidal), error_heuristic_default<error_metric_absolute>(error_metric_absolute()), 32);
    // finite rest
    const float v1 = integrate(integrator_fubini<1>(adaptive, monte_carlo(256, 1)), Slope3(), range_primary<3>());
    std::printf("fubini<1>(adaptive, monte_carlo(256)) finite: %.4f should be close to 0.5\n", v1); bad += std::fabs(v1 - 0.5f) > 0.03f;
    std::vector<float> bins(8, 0.0f);
    integrate(integrator_fubini<1>(monte_carlo_per_bin_parallel(256, 3), monte_carlo(16, 1)), bins, Slope3(), range_primary<3>());
    for (std::size_t i = 0; i < bins.size(); ++i) { const float want = 1.0f - (float(i) + 0.5f)/8.0f; std::printf(" bin %zu: %.4f (%.4f)", i, bins[i], want); bad += std::fabs(bins[i] - want) > 0.05f; }
    std::printf("\n");
    // infinite rest
    const float v2 = integrate(integrator_fubini<1>(adaptive, monte_carlo(4096, 2)), Decay(), range_primary_infinite<float>());
    std::printf("fubini<1>(adaptive, monte_carlo(4096)) infinite: %.3f should be close to 3 (heavy-tailed)\n", v2); bad += std::fabs(v2 - 3.0f) > 0.6f;
    // crespo2021_infinite: control variates over the first two dimensions, residual over the whole path
    tensor<float,2> img({16,16}, -1.0f), ref({16,16}, 0.0f);
    integrate(integrator_crespo2021_infinite<2>(256, 16, 64, 5), img, img.resolution(), Path(), range_primary_infinite<float>());
    integrate(monte_carlo_per_bin_parallel(16384, 9), ref, ref.resolution(), Path(), range_primary_infinite<float>());
    double m = 0, e = 0; for (std::size_t i = 0; i < img.size(); ++i) { m += img.raw_data()[i]; const double d = img.raw_data()[i] - ref.raw_data()[i]; e += d*d; }
    m /= img.size(); e = std::sqrt(e / img.size());
    std::printf("crespo2021_infinite<2>(256,16,64): mean of bins %.4f should be close to 1.013, rms error vs 16384-spp reference %.4f\n", m, e);
    bad += std::fabs(m - 1.013) > 0.04 || e > 0.35;
    std::vector<float> ten(10, 0.0f);
    integrate(integrator_crespo2021_infinite<4>(1280, 8, 256, 7), ten, Decay(), range_primary_infinite<float>());      // crespo21.cc:31
    for (float x : ten) std::printf("%.3f ", x);
    std::printf("<- crespo2021_infinite<4>(1280,8,256), 10 bins of the geometric series (3.x for x0 < 0.75, 0 beyond)\n");
    bad += !(ten[9] == 0.0f && ten[8] == 0.0f && ten[0] > 1.5f && ten[0] < 6.0f);
    return bad;
}
