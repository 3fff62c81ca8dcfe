// Newton-Cotes rule arithmetic on the device, replacing the reference's
//   Trapezoidal / Simpson / Boole  (src/newton-cotes/rules.h:9-44, 59-100, 251-293): operator() (weights),
//                                   coefficients (monomial form), at (Horner), subrange (antiderivative difference)
//   Nested<H,L>::low / error        (src/nested/nested.h:17-33)
//   error_metric_absolute/relative  (src/nested/error-metric.h:10-41)
// The reference mixes float and double inside these expressions (double literals in the weights and in the
// antiderivative, integer literals in the coefficients — SURVEY.md App. A #11) and rounds back to float at every
// return.  The bit-exact modes (greedy refinement order, region->bin integration, control-variate replay) only work
// if every one of those roundings happens here too, so each operation is spelled with an explicit round-to-nearest
// intrinsic: the results do not depend on whether the including TU is compiled with --fmad=true or false.
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

// quadrature weights, rules.h:14 / :64 / :256
template<int S> __device__ __forceinline__ float apply(const float* p);
template<> __device__ __forceinline__ float apply<2>(const float* p) { return d2f(dd(double(fa(p[0], p[1])), 2.0)); }
template<> __device__ __forceinline__ float apply<3>(const float* p) { return d2f(dd(da(da(double(p[0]), dm(4.0, double(p[1]))), double(p[2])), 6.0)); }
template<> __device__ __forceinline__ float apply<5>(const float* p) {
    double s = dm(7.0, double(p[0]));
    s = da(s, dm(32.0, double(p[1]))); s = da(s, dm(12.0, double(p[2]))); s = da(s, dm(32.0, double(p[3]))); s = da(s, dm(7.0, double(p[4])));
    return d2f(dd(s, 90.0));
}

// monomial coefficients, rules.h:27-32 / :77-83 / :272-280 (integer literals -> float arithmetic, left to right)
template<int S> __device__ __forceinline__ void coefficients(const float* p, float* c);
template<> __device__ __forceinline__ void coefficients<2>(const float* p, float* c) { c[0] = p[0]; c[1] = fs(p[1], p[0]); }
template<> __device__ __forceinline__ void coefficients<3>(const float* p, float* c) {
    c[0] = p[0];
    c[1] = fs(fa(fm(-3.0f, p[0]), fm(4.0f, p[1])), p[2]);
    c[2] = fa(fs(fm(2.0f, p[0]), fm(4.0f, p[1])), fm(2.0f, p[2]));
}
template<> __device__ __forceinline__ void coefficients<5>(const float* p, float* c) {
    c[0] = p[0];
    c[1] = fs(fa(fs(fa(fd(fm(-25.0f, p[0]), 3.0f), fm(16.0f, p[1])), fm(12.0f, p[2])), fd(fm(16.0f, p[3]), 3.0f)), p[4]);
    c[2] = fa(fs(fa(fs(fd(fm(70.0f, p[0]), 3.0f), fd(fm(208.0f, p[1]), 3.0f)), fm(76.0f, p[2])), fd(fm(112.0f, p[3]), 3.0f)), fd(fm(22.0f, p[4]), 3.0f));
    c[3] = fs(fa(fs(fa(fd(fm(-80.0f, p[0]), 3.0f), fm(96.0f, p[1])), fm(128.0f, p[2])), fd(fm(224.0f, p[3]), 3.0f)), fm(16.0f, p[4]));
    c[4] = fa(fs(fa(fs(fd(fm(32.0f, p[0]), 3.0f), fd(fm(128.0f, p[1]), 3.0f)), fm(64.0f, p[2])), fd(fm(128.0f, p[3]), 3.0f)), fd(fm(32.0f, p[4]), 3.0f));
}

// Horner evaluation in float, rules.h:35-38 / :86-89 / :283-286
template<int S> __device__ __forceinline__ float at(float t, const float* p) {
    float c[S]; coefficients<S>(p, c);
    float v = c[S - 1];
#pragma unroll
    for (int k = S - 2; k >= 0; --k) v = fa(fm(v, t), c[k]);
    return v;
}

// antiderivative at x of the interpolating polynomial: (((c4*x/5.0 + c3/4.0)*x + c2/3.0)*x + c1/2.0)*x + c0)*x  — c*x is a float
// product, the division by the double literal promotes the rest (rules.h:41-44 / :97-100 / :289-293)
template<int S> __device__ __forceinline__ double antiderivative(const float* c, float x) {
    double u = dd(double(fm(c[S - 1], x)), double(S));
#pragma unroll
    for (int k = S - 2; k >= 1; --k) { u = da(u, dd(double(c[k]), double(k + 1))); u = dm(u, double(x)); }
    u = da(u, double(c[0]));
    return dm(u, double(x));
}
template<int S> __device__ __forceinline__ float subrange(float a, float b, const float* p) {
    float c[S]; coefficients<S>(p, c);
    return d2f(ds(antiderivative<S>(c, b), antiderivative<S>(c, a)));
}

// nested.h:17-23
template<int SH, int SL> __device__ __forceinline__ float low(const float* p) {
    float q[SL];
#pragma unroll
    for (int i = 0; i < SL; ++i) q[i] = p[i * (SH - 1) / (SL - 1)];
    return apply<SL>(q);
}

// error-metric.h:10-13 / :30-37
__device__ __forceinline__ float metric(bool relative, float a, float b) {
    const float diff = fabsf(fs(b, a));
    if (!relative) return diff;
    const float m = fmaxf(fabsf(a), fabsf(b));
    if (double(m) < 1.e-37) return diff;
    return fd(diff, m);
}
// nested.h:31-33
template<int SH, int SL> __device__ __forceinline__ float line_error(bool relative, const float* p) {
    return metric(relative, apply<SH>(p), low<SH, SL>(p));
}

// range.h:45-53
__device__ __forceinline__ float pos_in_range(float lo, float hi, float p) { return (lo >= hi) ? lo : fd(fs(p, lo), fs(hi, lo)); }

__host__ __device__ constexpr int ipow(int s, int d) { return d <= 0 ? 1 : s * ipow(s, d - 1); }

}}}} // namespace viltrum::b200::device::rules
