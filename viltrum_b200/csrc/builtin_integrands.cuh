// Device twins of the synthetic integrands of SURVEY.md §8(d) / Appendix D (shade4<K>, shade5<K>, smooth_edge2,
// walk) plus the small analytic ones the reference's own examples use (x^2+y^2: main/doc/montecarlo-2d.cc:10;
// (x+y<1): main/compilation-tests/array-parameter.cc:11-14; geometric-series walk: main/doc/montecarlo-infd.cc:8-22).
// Transcendental-free fp32, so a --fmad=false build evaluates them to the same bits as a CPU build with
// -ffp-contract=off.  Written from the formulas in SURVEY.md; nothing here comes from oracle/.
#pragma once
#include <array>
#include <viltrum_b200/device/f32x2.cuh>

namespace viltrum { namespace b200 { namespace builtin {

struct X2Y2 { __host__ __device__ float operator()(const std::array<float,2>& x) const { return x[0]*x[0] + x[1]*x[1]; } };
struct Ind2 { __host__ __device__ float operator()(const std::array<float,2>& x) const { return ((x[0]+x[1])<1.0f)?1.0f:0.0f; } };
struct Cubic1 { __host__ __device__ float operator()(const std::array<float,1>& x) const { return (4.0f*x[0]*x[0]-1.0f)*x[0] + 0.25f; } };
struct Poly3 { __host__ __device__ float operator()(const std::array<float,3>& x) const { return x[0]*x[1] + x[1]*x[2]*x[2] + 0.5f; } };

// Generic over the scalar type (float, or viltrum::b200::f32x2 = two samples per call on the packed FP32 pipe, device/f32x2.cuh).
// For T = float the expression tree is the original one (mad(a,b,c) is literally a*b+c), so the exact build's bits are unchanged.
template<int K> struct Shade4 {
    template<class T> __host__ __device__ T operator()(const std::array<T,4>& x) const {
        using viltrum::b200::mad; using viltrum::b200::indicator;
        const T a = x[0]-.5f, b = x[1]-.5f;
        const T edge = .55f+.35f*(a*a-b*b)+.2f*a*b;
        const T vis = indicator(x[2]+.5f*x[3]<edge);
        const T t = x[2]*(1.0f-x[3]);
        T lobe = T(1.0f/float(K));
#pragma unroll
        for (int k=K-2;k>=0;--k) lobe = mad(lobe,t,1.0f/float(k+1));      // Horner, c_k = 1/(k+1)
        const T alb = .25f+.75f*x[0]*x[1];
        return vis*lobe*alb;
    }
};
template<int K> struct Shade5 {
    template<class T> __host__ __device__ T operator()(const std::array<T,5>& x) const {
        return Shade4<K>()(std::array<T,4>{x[0],x[1],x[2],x[3]})*(.5f+x[4]);
    }
};
struct SmoothEdge2 {
    __host__ __device__ float operator()(const std::array<float,2>& p) const {
        const float x = p[0], y = p[1];
        const float s = .5f+8.0f*x*(1.0f-x)*y*(1.0f-y)*(1.0f-2.0f*(x-y)*(x-y));
        const float dx = x-.45f, dy = y-.55f;
        return s+((dx*dx+dy*dy<.09f)?.75f:0.0f);
    }
};
// ---- double-precision family (Range<double,DIM>): same shapes, double constants --------------------------------------------
struct X2Y2d { __host__ __device__ double operator()(const std::array<double,2>& x) const { return x[0]*x[0] + x[1]*x[1]; } };
struct Ind2d { __host__ __device__ double operator()(const std::array<double,2>& x) const { return ((x[0]+x[1])<1.0)?1.0:0.0; } };
struct Cubic1d { __host__ __device__ double operator()(const std::array<double,1>& x) const { return (4.0*x[0]*x[0]-1.0)*x[0] + 0.25; } };
struct Poly3d { __host__ __device__ double operator()(const std::array<double,3>& x) const { return x[0]*x[1] + x[1]*x[2]*x[2] + 0.5; } };
struct SmoothEdge2d {
    __host__ __device__ double operator()(const std::array<double,2>& p) const {
        const double x = p[0], y = p[1];
        const double s = 0.5+8.0*x*(1.0-x)*y*(1.0-y)*(1.0-2.0*(x-y)*(x-y));
        const double dx = x-0.45, dy = y-0.55;
        return s+((dx*dx+dy*dy<0.09)?0.75:0.0);
    }
};
template<int K> struct Shade4d {
    __host__ __device__ double operator()(const std::array<double,4>& x) const {
        const double a = x[0]-0.5, b = x[1]-0.5;
        const double edge = 0.55+0.35*(a*a-b*b)+0.2*a*b;
        const double vis = (x[2]+0.5*x[3]<edge)?1.0:0.0;
        const double t = x[2]*(1.0-x[3]);
        double lobe = 1.0/double(K);
        for (int k=K-2;k>=0;--k) lobe = lobe*t+1.0/double(k+1);
        const double alb = 0.25+0.75*x[0]*x[1];
        return vis*lobe*alb;
    }
};

struct Walk {
    template<typename Seq> __host__ __device__ float operator()(const Seq& seq) const {
        auto it = seq.begin(); const float px = *it; ++it; const float py = *it; ++it;
        const float alb = .4f+.5f*(4.0f*px*(1.0f-px))*(.25f+.75f*py);
        float pos = .5f, L = 0.0f;
        while (true) { const float u = *it; ++it; if (u>=alb) break;
                       const float s = *it; ++it; pos = .5f*pos+.5f*s; L += .25f+pos*pos; }
        return L;
    }
    // the same path as a state machine (include/viltrum_b200/device/walk.cuh, wavefront kernel): identical arithmetic
    struct State { float alb, pos, L; };
    static constexpr int elements_begin = 2, elements_step = 2;      // whole Philox blocks: walk_block_kernel
    template<typename It> __host__ __device__ State begin(It& it) const {
        const float px = *it; ++it; const float py = *it; ++it;
        return State{.4f+.5f*(4.0f*px*(1.0f-px))*(.25f+.75f*py), .5f, 0.0f};
    }
    template<typename It> __host__ __device__ bool step(State& st, It& it) const {
        const float u = *it; ++it; if (u>=st.alb) return false;
        const float s = *it; ++it; st.pos = .5f*st.pos+.5f*s; st.L += .25f+st.pos*st.pos;
        return true;
    }
    __host__ __device__ float end(const State& st) const { return st.L; }
};
// same integrand with the state-machine form but WITHOUT the element counts: keeps walk_wavefront_kernel (batched refill through the
// general iterator) measurable and tested
struct WalkSteps {
    template<typename Seq> __host__ __device__ float operator()(const Seq& seq) const { return Walk()(seq); }
    using State = Walk::State;
    template<typename It> __host__ __device__ State begin(It& it) const { return Walk().begin(it); }
    template<typename It> __host__ __device__ bool step(State& st, It& it) const { return Walk().step(st, it); }
    __host__ __device__ float end(const State& st) const { return st.L; }
};
// same integrand WITHOUT the state-machine form: keeps the generic per-lane kernel measurable and tested
struct WalkPlain {
    template<typename Seq> __host__ __device__ float operator()(const Seq& seq) const { return Walk()(seq); }
};
struct Decay {
    template<typename Seq> __host__ __device__ float operator()(const Seq& seq) const {
        auto x = seq.begin(); float sum = 0.0f, term = 1.0f;
        while ((*x) < 0.75f) { ++x; term *= 2.0f*(*x); ++x; sum += term; }
        return sum;
    }
    // the same series as a state machine that reads nothing in begin() and two elements per round (walk_block_kernel)
    struct State { float sum, term; };
    static constexpr int elements_begin = 0, elements_step = 2;
    template<typename It> __host__ __device__ State begin(It&) const { return State{0.0f, 1.0f}; }
    template<typename It> __host__ __device__ bool step(State& st, It& x) const {
        if (!((*x) < 0.75f)) return false;
        ++x; st.term *= 2.0f*(*x); ++x; st.sum += st.term;
        return true;
    }
    __host__ __device__ float end(const State& st) const { return st.sum; }
};
// operator()(seq) only: the generic per-lane kernel
struct DecayPlain {
    template<typename Seq> __host__ __device__ float operator()(const Seq& seq) const { return Decay()(seq); }
};

}}} // namespace viltrum::b200::builtin
