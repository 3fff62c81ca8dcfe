// Newton-Cotes rule arithmetic on the device, replacing the reference's
//   Trapezoidal / Simpson / Boole  (src/newton-cotes/rules.h:9-44, 59-100, 251-293): operator() (weights),
//                                   coefficients (monomial form), at (Horner), subrange (antiderivative difference)
//   Nested<H,L>::low / error        (src/nested/nested.h:17-33)
//   error_metric_absolute/relative  (src/nested/error-metric.h:10-41)
// The reference mixes its Float type and double inside these expressions (double literals in the weights and in the
// antiderivative, integer literals in the coefficients — SURVEY.md App. A #11) and rounds back to Float at every
// return.  The bit-exact modes (greedy refinement order, region->bin integration, control-variate replay) only work
// if every one of those roundings happens here too, so each operation is spelled with an explicit round-to-nearest
// intrinsic: the results do not depend on whether the including TU is compiled with --fmad=true or false.
// Everything is a template over T = Float = value_type (float, or double for Range<double,DIM>: then the promotions
// are no-ops and the whole expression is evaluated in double, exactly as upstream).
#pragma once
#include <cuda_runtime.h>

namespace viltrum { namespace b200 { namespace device { namespace rules {

__device__ __forceinline__ float  fm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float  fa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float  fs(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float  fd(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double dm(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double da(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double ds(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dd(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float  d2f(double a) { return __double2float_rn(a); }

// a / K for the small integer constants of the rules (2, 3, 4, 5, 6, 90), correctly rounded without the division subroutine:
// powers of two are exact multiplications; otherwise Markstein's sequence q0 = RN(a*y), r = a - K*q0 (exact in one FMA),
// q = RN(q0 + r*y) with y = RN(1/K), which is the correctly rounded quotient for every K here (no all-ones significand) as long as
// nothing under- or overflows — checked against a/K on 4e8 random doubles (profiles/exp/divtest.c) and guarded by magnitude, with
// the true division as the fallback (zeros, infinities, NaNs, denormal neighbourhoods).
template<int K> __device__ __forceinline__ double divc(double a) {
    if constexpr (K == 2) return __dmul_rn(a, 0.5);
    else if constexpr (K == 4) return __dmul_rn(a, 0.25);
    else {
        constexpr double y = 1.0 / double(K);
        const double m = fabs(a);
        if (!(m >= 1e290) && !(m > 0.0 && m <= 1e-290)) {       // zeros and NaNs take the fast path too (q0 = +-0 / NaN is already the answer)
            const double q0 = __dmul_rn(a, y);
            const double r = __fma_rn(-double(K), q0, a);
            const double q = __fma_rn(r, y, q0);
            return r == 0.0 ? q0 : q;                            // keeps the sign of a zero quotient
        }
        return __ddiv_rn(a, double(K));
    }
}

// arithmetic in T with explicit rounding
__device__ __forceinline__ float  mul(float a, float b) { return fm(a, b); }
__device__ __forceinline__ float  add(float a, float b) { return fa(a, b); }
__device__ __forceinline__ float  sub(float a, float b) { return fs(a, b); }
__device__ __forceinline__ float  quo(float a, float b) { return fd(a, b); }
__device__ __forceinline__ double mul(double a, double b) { return dm(a, b); }
__device__ __forceinline__ double add(double a, double b) { return da(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return ds(a, b); }
__device__ __forceinline__ double quo(double a, double b) { return dd(a, b); }
template<class T> __device__ __forceinline__ T from_double(double a);
template<> __device__ __forceinline__ float  from_double<float>(double a) { return d2f(a); }
template<> __device__ __forceinline__ double from_double<double>(double a) { return a; }
__device__ __forceinline__ float  absv(float a) { return fabsf(a); }
__device__ __forceinline__ double absv(double a) { return fabs(a); }
__device__ __forceinline__ float  maxv(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double maxv(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float  minv(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double minv(double a, double b) { return fmin(a, b); }

// quadrature weights, rules.h:14 / :64 / :256
template<int S, class T> __device__ __forceinline__ T apply(const T* p) {
    if constexpr (S == 2) return from_double<T>(divc<2>(double(add(p[0], p[1]))));
    else if constexpr (S == 3) return from_double<T>(divc<6>(da(da(double(p[0]), dm(4.0, double(p[1]))), double(p[2]))));
    else {
        double s = dm(7.0, double(p[0]));
        s = da(s, dm(32.0, double(p[1]))); s = da(s, dm(12.0, double(p[2]))); s = da(s, dm(32.0, double(p[3]))); s = da(s, dm(7.0, double(p[4])));
        return from_double<T>(divc<90>(s));
    }
}

// monomial coefficients, rules.h:27-32 / :77-83 / :272-280 (integer literals -> arithmetic in T, left to right)
template<int S, class T> __device__ __forceinline__ void coefficients(const T* p, T* c) {
    if constexpr (S == 2) { c[0] = p[0]; c[1] = sub(p[1], p[0]); }
    else if constexpr (S == 3) {
        c[0] = p[0];
        c[1] = sub(add(mul(T(-3), p[0]), mul(T(4), p[1])), p[2]);
        c[2] = add(sub(mul(T(2), p[0]), mul(T(4), p[1])), mul(T(2), p[2]));
    } else {
        c[0] = p[0];
        c[1] = sub(add(sub(add(quo(mul(T(-25), p[0]), T(3)), mul(T(16), p[1])), mul(T(12), p[2])), quo(mul(T(16), p[3]), T(3))), p[4]);
        c[2] = add(sub(add(sub(quo(mul(T(70), p[0]), T(3)), quo(mul(T(208), p[1]), T(3))), mul(T(76), p[2])), quo(mul(T(112), p[3]), T(3))), quo(mul(T(22), p[4]), T(3)));
        c[3] = sub(add(sub(add(quo(mul(T(-80), p[0]), T(3)), mul(T(96), p[1])), mul(T(128), p[2])), quo(mul(T(224), p[3]), T(3))), mul(T(16), p[4]));
        c[4] = add(sub(add(sub(quo(mul(T(32), p[0]), T(3)), quo(mul(T(128), p[1]), T(3))), mul(T(64), p[2])), quo(mul(T(128), p[3]), T(3))), quo(mul(T(32), p[4]), T(3)));
    }
}

// Horner evaluation in T, rules.h:35-38 / :86-89 / :283-286
template<int S, class T> __device__ __forceinline__ T at(T t, const T* p) {
    T c[S]; coefficients<S, T>(p, c);
    T v = c[S - 1];
#pragma unroll
    for (int k = S - 2; k >= 0; --k) v = add(mul(v, t), c[k]);
    return v;
}

// antiderivative at x of the interpolating polynomial: (((c4*x/5.0 + c3/4.0)*x + c2/3.0)*x + c1/2.0)*x + c0)*x  — c*x is a product in T,
// the division by the double literal promotes the rest (rules.h:41-44 / :97-100 / :289-293)
template<int S, int K> struct AntiStep {       // the Horner steps k = K .. 1, unrolled at compile time so that every divisor is a constant
    template<class T> __device__ __forceinline__ static double run(double u, const T* c, T x) {
        u = da(u, divc<K + 1>(double(c[K]))); u = dm(u, double(x));
        return AntiStep<S, K - 1>::run(u, c, x);
    }
};
template<int S> struct AntiStep<S, 0> { template<class T> __device__ __forceinline__ static double run(double u, const T*, T) { return u; } };
template<int S, class T> __device__ __forceinline__ double antiderivative(const T* c, T x) {
    double u = divc<S>(double(mul(c[S - 1], x)));
    u = AntiStep<S, S - 2>::run(u, c, x);
    u = da(u, double(c[0]));
    return dm(u, double(x));
}
template<int S, class T> __device__ __forceinline__ T subrange(T a, T b, const T* p) {
    T c[S]; coefficients<S, T>(p, c);
    return from_double<T>(ds(antiderivative<S, T>(c, b), antiderivative<S, T>(c, a)));
}

// Simpson::pdf_points + pdf_integral_subrange (rules.h:104-154): |p| (NormDefault), shifted down by the minimum of its parabola where
// that minimum is negative and inside (0,1), integrated over [t0,t1]
template<class T> __device__ __forceinline__ T simpson_pdf_integral_subrange(T t0, T t1, const T* p) {
    T q[3] = {absv(p[0]), absv(p[1]), absv(p[2])};
    T cs[3]; coefficients<3, T>(q, cs);
    T ymin = T(0);
    if (cs[2] > T(0)) {
        const T tmin = from_double<T>(dd(double(-cs[1]), dm(2.0, double(cs[2]))));       // -cs[1]/(2.0*cs[2]), double literal (rules.h:123)
        if (tmin > T(0) && tmin < T(1)) {
            const T ytmin = add(mul(add(mul(cs[2], tmin), cs[1]), tmin), cs[0]);
            if (ytmin < ymin) ymin = ytmin;
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) q[i] = sub(q[i], ymin);
    return subrange<3, T>(t0, t1, q);
}

// nested.h:17-23
template<int SH, int SL, class T> __device__ __forceinline__ T low(const T* p) {
    T q[SL];
#pragma unroll
    for (int i = 0; i < SL; ++i) q[i] = p[i * (SH - 1) / (SL - 1)];
    return apply<SL, T>(q);
}

// error-metric.h:10-13 / :30-37
template<class T> __device__ __forceinline__ T metric(bool relative, T a, T b) {
    const T diff = absv(sub(b, a));
    if (!relative) return diff;
    const T m = maxv(absv(a), absv(b));
    if (double(m) < 1.e-37) return diff;
    return quo(diff, m);
}
// nested.h:31-33
template<int SH, int SL, class T> __device__ __forceinline__ T line_error(bool relative, const T* p) {
    return metric<T>(relative, apply<SH, T>(p), low<SH, SL, T>(p));
}

// range.h:45-53
template<class T> __device__ __forceinline__ T pos_in_range(T lo, T hi, T p) { return (lo >= hi) ? lo : quo(sub(p, lo), sub(hi, lo)); }

__host__ __device__ constexpr int ipow(int s, int d) { return d <= 0 ? 1 : s * ipow(s, d - 1); }

}}}} // namespace viltrum::b200::device::rules
