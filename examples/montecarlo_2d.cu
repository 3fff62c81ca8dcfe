// BASELINE.json config 1 — the README example (reference main/doc/montecarlo-2d.cc, README.md:60-100):
// monte_carlo(8192), f(x,y) = x^2 + y^2 over [0,1]^2 into 10 bins, through an accessor and through std::vector<float>.
#include <viltrum_b200/viltrum.h>
#include <cstdio>

struct X2Y2 { __host__ __device__ float operator()(const std::array<float,2>& x) const { return x[0]*x[0] + x[1]*x[1]; } };

int main() {
    float sol[10] = {0};      // the reference accumulates ('+=') into whatever the bins hold: start from zero
    auto sol_access = [&sol] (const std::array<std::size_t,1>& pos) -> float& { return sol[pos[0]]; };
    auto range = viltrum::range(std::array<float,2>{0.0f,0.0f}, std::array<float,2>{1.0f,1.0f});
    viltrum::integrate(viltrum::monte_carlo(8192, 0), sol_access, std::array<std::size_t,1>{10}, X2Y2(), range);
    std::vector<float> sol_vec(10, 0.0f);
    viltrum::integrate(viltrum::monte_carlo(1u << 22, 1), sol_vec, X2Y2(), range);
    // two-argument integrand f(x,y) (integrate.h:26-31)
    std::vector<float> sol_xy(10, 0.0f);
    auto fxy = [] __host__ __device__ (float x, float y) { return x*x + y*y; };
    viltrum::integrate(viltrum::monte_carlo(1u << 22, 2), sol_xy, fxy, range);
    int bad = 0;
    for (int i = 0; i < 10; ++i) {
        const float analytic = (3.0f*i*i + 3.0f*i + 1.0f)/300.0f + 1.0f/3.0f;
        std::printf("Bin %d: %.5f  %.5f  %.5f  (analytic %.5f)\n", i, sol[i], sol_vec[i], sol_xy[i], analytic);
        if (std::fabs(sol[i]-analytic) > 0.15f || std::fabs(sol_vec[i]-analytic) > 0.006f || std::fabs(sol_xy[i]-analytic) > 0.006f) ++bad;
    }
    // nested std::vector bins (integrate.h:139-167): a ragged 4 x (<=5) container is squared up to 4 x 5; bin (i,j) of x^2+y^2 over the unit square
    std::vector<std::vector<float>> grid(4); grid[2].resize(5);
    viltrum::LoggerNull lognull;
    viltrum::integrate(viltrum::monte_carlo_per_bin_parallel(4096, 5), grid, X2Y2(), range, lognull);
    for (std::size_t i = 0; i < 4; ++i) for (std::size_t j = 0; j < 5; ++j) {
        const double x0 = i/4.0, x1 = (i+1)/4.0, y0 = j/5.0, y1 = (j+1)/5.0;
        const double want = 20.0*((x1*x1*x1-x0*x0*x0)/3.0*(y1-y0) + (y1*y1*y1-y0*y0*y0)/3.0*(x1-x0));      // nbins x the bin's integral = the mean of f over the bin
        if (grid[i].size() != 5 || std::fabs(grid[i][j] - want) > 0.02*want + 1e-4) ++bad;
    }
    std::printf("nested vector bins: grid[3][4] = %.5f\n", grid[3][4]);
    float single = viltrum::integrate(viltrum::monte_carlo(1u << 20, 3), X2Y2(), range);
    std::printf("single value: %.5f should be close to %.5f\n", single, 2.0f/3.0f);
    if (std::fabs(single - 2.0f/3.0f) > 0.004f) ++bad;
    return bad ? 1 : 0;
}
