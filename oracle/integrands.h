// TEST INFRASTRUCTURE — not product code.
//
// Host-only synthetic integrands shared by the two CPU checkers under oracle/:
//   * ref_harness_*.cpp  -> oracle/_ref/libviltrum_ref.so   (the UNMODIFIED reference, compiled from
//                           /root/reference where it lies)
//   * oracle.cpp         -> oracle/liboracle.so             (plain restatement of the algorithms)
// The shapes are the ones SURVEY.md §8(d) / Appendix D define (shade4<K>, shade5<K>, smooth_edge2,
// walk) plus a few small analytic ones for known-answer tests.  All are transcendental-free fp32 so
// that CPU (-ffp-contract=off) and GPU (--fmad=false) evaluate them to identical bits.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
// anything under oracle/.  The product (viltrum_b200/, include/) never includes this file; the
// device-side twins live in viltrum_b200/csrc/builtin_integrands.cuh.
#pragma once
#include <array>
#include <cstddef>

namespace vo {

// f(x,y) = x^2 + y^2  — README example (reference main/doc/montecarlo-2d.cc:10)
struct X2Y2 {
    static constexpr int dim = 2;
    float operator()(const std::array<float,2>& x) const { return x[0]*x[0] + x[1]*x[1]; }
};

// indicator (x+y<1) — reference main/compilation-tests/array-parameter.cc:11-14 (float-valued here)
struct Ind2 {
    static constexpr int dim = 2;
    float operator()(const std::array<float,2>& x) const { return ((x[0]+x[1])<1.0f)?1.0f:0.0f; }
};

// 1-D cubic, exactly integrable by Simpson/Boole: 4x^3 - x + 0.25
struct Cubic1 {
    static constexpr int dim = 1;
    float operator()(const std::array<float,1>& x) const { return (4.0f*x[0]*x[0]-1.0f)*x[0] + 0.25f; }
};

// 3-D smooth polynomial: x*y + y*z*z + 0.5
struct Poly3 {
    static constexpr int dim = 3;
    float operator()(const std::array<float,3>& x) const { return x[0]*x[1] + x[1]*x[2]*x[2] + 0.5f; }
};

// shade4<K>  (20 + 2(K-1) flops)
template<int K> struct Shade4 {
    static constexpr int dim = 4;
    float operator()(const std::array<float,4>& x) const {
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

template<int K> struct Shade5 {
    static constexpr int dim = 5;
    float operator()(const std::array<float,5>& x) const {
        return Shade4<K>()(std::array<float,4>{x[0],x[1],x[2],x[3]})*(.5f+x[4]);
    }
};

struct SmoothEdge2 {
    static constexpr int dim = 2;
    float operator()(const std::array<float,2>& p) const {
        float x=p[0], y=p[1];
        float s=.5f+8.0f*x*(1.0f-x)*y*(1.0f-y)*(1.0f-2.0f*(x-y)*(x-y));
        float dx=x-.45f, dy=y-.55f;
        return s+((dx*dx+dy*dy<.09f)?.75f:0.0f);
    }
};

// ---- double-precision family (Range<double,DIM>, north_star: Newton-Cotes within 1e-12 in fp64) ------------------------------
// Same shapes with double constants; these are integrands in their own right (0.55 != 0.55f), named like the float ones and
// selected by the *_f64 entry points.
template<typename T> struct X2Y2T { static constexpr int dim = 2; T operator()(const std::array<T,2>& x) const { return x[0]*x[0] + x[1]*x[1]; } };
template<typename T> struct Ind2T { static constexpr int dim = 2; T operator()(const std::array<T,2>& x) const { return ((x[0]+x[1])<T(1))?T(1):T(0); } };
template<typename T> struct Cubic1T { static constexpr int dim = 1; T operator()(const std::array<T,1>& x) const { return (T(4)*x[0]*x[0]-T(1))*x[0] + T(0.25); } };
template<typename T> struct Poly3T { static constexpr int dim = 3; T operator()(const std::array<T,3>& x) const { return x[0]*x[1] + x[1]*x[2]*x[2] + T(0.5); } };
template<typename T> struct SmoothEdge2T {
    static constexpr int dim = 2;
    T operator()(const std::array<T,2>& p) const {
        T x=p[0], y=p[1];
        T s=T(0.5)+T(8)*x*(T(1)-x)*y*(T(1)-y)*(T(1)-T(2)*(x-y)*(x-y));
        T dx=x-T(0.45), dy=y-T(0.55);
        return s+((dx*dx+dy*dy<T(0.09))?T(0.75):T(0));
    }
};
template<typename T, int K> struct Shade4T {
    static constexpr int dim = 4;
    T operator()(const std::array<T,4>& x) const {
        T a=x[0]-T(0.5), b=x[1]-T(0.5);
        T edge=T(0.55)+T(0.35)*(a*a-b*b)+T(0.2)*a*b;
        T vis=(x[2]+T(0.5)*x[3]<edge)?T(1):T(0);
        T t=x[2]*(T(1)-x[3]);
        T lobe=T(1)/T(K);
        for (int k=K-2;k>=0;--k) lobe=lobe*t+T(1)/T(k+1);
        T alb=T(0.25)+T(0.75)*x[0]*x[1];
        return vis*lobe*alb;
    }
};

// Infinite-dimensional random walk with Russian roulette.
struct Walk {
    template<typename Seq> float operator()(const Seq& seq) const {
        auto it=seq.begin(); float px=*it; ++it; float py=*it; ++it;
        float alb=.4f+.5f*(4.0f*px*(1.0f-px))*(.25f+.75f*py);
        float pos=.5f, L=0.0f;
        while (true) { float u=*it; ++it; if (u>=alb) break;
                       float s=*it; ++it; pos=.5f*pos+.5f*s; L+=.25f+pos*pos; }
        return L;
    }
};

// Geometric series walk from reference main/doc/montecarlo-infd.cc:8-22 (decay 0.75 -> integral 3).
struct Decay {
    template<typename Seq> float operator()(const Seq& seq) const {
        auto x = seq.begin(); float sum=0.0f, term=1.0f;
        while ((*x) < 0.75f) { ++x; term *= 2.0f*(*x); ++x; sum += term; }
        return sum;
    }
};

} // namespace vo
