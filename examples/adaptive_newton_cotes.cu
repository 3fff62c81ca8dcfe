// BASELINE.json config 3 — nested Newton-Cotes (Boole/Simpson) adaptive refinement with the greedy heap
// (SURVEY.md §3.3).  Built as an EXACT twin (--fmad=false): the bins written to argv[3] are bit-identical to the
// reference's integrator_adaptive_iterations on the same integrand (tests/test_gpu_examples.py compares them).
#include <viltrum_b200/viltrum.h>
#include <cstdio>
#include <cstdlib>

struct SmoothEdge2 {
    __host__ __device__ float operator()(const std::array<float,2>& p) const {
        float x=p[0], y=p[1];
        float s=.5f+8.0f*x*(1.0f-x)*y*(1.0f-y)*(1.0f-2.0f*(x-y)*(x-y));
        float dx=x-.45f, dy=y-.55f;
        return s+((dx*dx+dy*dy<.09f)?.75f:0.0f);
    }
};
struct X2Y2d { __host__ __device__ double operator()(const std::array<double,2>& x) const { return x[0]*x[0] + x[1]*x[1]; } };
struct CountLogger : viltrum::LoggerNull {           // the reference hands the region list to Logger::log (integrator-region-based.h:19)
    std::size_t* regions;
    template<typename Data> void log(const Data& d) { *regions = d.size(); }
};

int main(int argc, char** argv) {
    using namespace viltrum;
    const std::size_t w = argc > 1 ? std::atoi(argv[1]) : 64, iterations = argc > 2 ? std::atoi(argv[2]) : 20000;
    tensor<float,2> img({w,w}, 0.0f);
    std::size_t nregions = 0; CountLogger logger; logger.regions = &nregions;
    integrate(integrator_adaptive_iterations(nested(boole,simpson), error_heuristic_size(error_metric_relative(),1.e-5), iterations),
              img, img.resolution(), SmoothEdge2(), range_primary<2>(), logger);
    std::vector<float> fixed(16, 0.0f);
    integrate(integrator_newton_cotes(simpson), fixed, [] __host__ __device__ (float x, float y) { return x*x + y*y; }, range_primary<2>());
    // Range<double,DIM> + double integrand: the fp64 Newton-Cotes path (1e-12 gate)
    std::vector<double> fixed64(16, 0.0);
    integrate(integrator_newton_cotes(boole), fixed64, X2Y2d(), range_primary<2,double>());
    std::printf("fp64 boole bin 0: %.15f (analytic %.15f)\n", fixed64[0], 1.0/768.0 + 1.0/3.0);
    // integrator_adaptive_tolerance (main/doc/integrators.cc:66-71): refine until every region's error estimate is below the tolerance
    tensor<float,2> tol({w,w}, 0.0f);
    integrate(integrator_adaptive_tolerance(nested(boole,simpson), error_heuristic_default(error_metric_absolute()), 1.e-7f), tol, tol.resolution(), SmoothEdge2(), range_primary<2>());
    double mt = 0; for (float v : tol.raw_data()) mt += v; mt /= tol.size();
    std::printf("adaptive_tolerance(1e-7): mean of bins %.6f\n", mt);
    if (argc > 4) { FILE* f = std::fopen(argv[4], "wb"); std::fwrite(tol.raw_data().data(), 4, tol.size(), f); std::fclose(f); }
    double m = 0; for (float v : img.raw_data()) m += v; m /= img.size();
    const double analytic = 0.5 + 2.0/9.0 - 2.0/45.0 + 0.75*3.14159265358979*0.09;
    std::printf("adaptive: %zu regions, mean of bins %.6f should be close to %.6f; simpson bin 0 %.6f\n", nregions, m, analytic, fixed[0]);
    if (argc > 3) { FILE* f = std::fopen(argv[3], "wb"); std::fwrite(img.raw_data().data(), 4, img.size(), f); std::fclose(f); }
    return (std::fabs(m-analytic) < 1e-3 && std::fabs(mt-analytic) < 1e-3 && nregions == iterations+1 && std::fabs(fixed64[0] - (1.0/768.0 + 1.0/3.0)) < 1e-13) ? 0 : 1;
}
