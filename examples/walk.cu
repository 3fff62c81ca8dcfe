// BASELINE.json config 5 — infinite-dimensional random walk with Russian roulette over range_primary_infinite
// (reference main/doc/montecarlo-infd.cc for the sequence protocol).
#include <viltrum_b200/viltrum.h>
#include <cstdio>
#include <cstdlib>

struct Walk {
    template<typename Seq> __host__ __device__ float operator()(const Seq& seq) const {
        auto it=seq.begin(); float px=*it; ++it; float py=*it; ++it;
        float alb=.4f+.5f*(4.0f*px*(1.0f-px))*(.25f+.75f*py);
        float pos=.5f, L=0.0f;
        while (true) { float u=*it; ++it; if (u>=alb) break;
                       float s=*it; ++it; pos=.5f*pos+.5f*s; L+=.25f+pos*pos; }
        return L;
    }
    // OPTIONAL (viltrum_b200 extension): the same path as a state machine with its element counts.  operator()(seq) alone runs on the generic
    // per-lane kernel; with begin/step/end the lanes of a warp refill independently, and with elements_begin/elements_step = 2/2 (or 0/2)
    // every lane is fed one whole Philox block per iteration (include/viltrum_b200/device/walk.cuh).  Same arithmetic -> same bins, bit for bit.
    struct State { float alb, pos, L; };
    static constexpr int elements_begin = 2, elements_step = 2;
    template<typename It> __host__ __device__ State begin(It& it) const {
        float px=*it; ++it; float py=*it; ++it;
        return State{.4f+.5f*(4.0f*px*(1.0f-px))*(.25f+.75f*py), .5f, 0.0f};
    }
    template<typename It> __host__ __device__ bool step(State& st, It& it) const {
        float u=*it; ++it; if (u>=st.alb) return false;
        float s=*it; ++it; st.pos=.5f*st.pos+.5f*s; st.L+=.25f+st.pos*st.pos;
        return true;
    }
    __host__ __device__ float end(const State& st) const { return st.L; }
};
struct Decay {                                       // reference main/doc/montecarlo-infd.cc:8-22
    float decaying_factor;
    template<typename Seq> __host__ __device__ float operator()(const Seq& seq) const {
        auto x = seq.begin(); float sum = 0.0f, term = 1.0f;
        while ((*x) < decaying_factor) { ++x; term *= 2.0f*(*x); ++x; sum += term; }
        return sum;
    }
};

int main(int argc, char** argv) {
    const std::size_t w = argc > 1 ? std::atoi(argv[1]) : 128, spp = argc > 2 ? std::atoi(argv[2]) : 256;
    viltrum::tensor<float,2> img({w,w}, 0.0f);
    viltrum::integrate(viltrum::monte_carlo_per_bin_parallel(spp, 0), img, img.resolution(), Walk(), viltrum::range_primary_infinite<float>());
    double m = 0; for (float v : img.raw_data()) m += v; m /= img.size();
    float sol = viltrum::integrate(viltrum::monte_carlo_per_bin_parallel(1u << 20, 1), Decay{0.75f}, viltrum::range_infinite<float>(0.0f, 1.0f));
    std::printf("walk: mean of bins %.5f should be close to 1.0133; decay integral %.5f should be close to %.5f\n", m, sol, 0.75f/(1.0f-0.75f));
    return (std::fabs(m-1.0133) < 5e-3 && std::fabs(sol-3.0f) < 0.03f) ? 0 : 1;
}
