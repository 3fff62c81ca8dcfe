// Packed pairs of floats (Blackwell FFMA2 / FMUL2 / FADD2: one instruction, two fp32 lanes per thread).
// The per-bin sampling kernel evaluates TWO samples per call when the integrand functor is generic over its scalar type, i.e.
//     template<class T> __host__ __device__ T operator()(const std::array<T,DIM>& x) const
// instantiates for T = float and for T = viltrum::b200::f32x2.  Arithmetic operators, comparisons and the helpers below
// (mad, indicator, select) are defined for both types, so one body serves both:
//     const T t = x[2]*(1.0f - x[3]);  T acc = T(c_n);  for (...) acc = mad(acc, t, c_k);  return indicator(x[0] < t)*acc;
// Half the issue slots for the FP32 work (measured: profiles/mc_packed_r1.txt); results per lane are those of the scalar code with
// every mad() fused.  Functors that only take std::array<float,DIM> keep working unchanged (scalar path).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstring>

#if defined(__CUDACC__)
#define VB200_F2 __host__ __device__ __forceinline__
#else
#define VB200_F2 inline
#endif

namespace viltrum { namespace b200 {

struct mask2 { bool lo, hi; };

struct alignas(8) f32x2 {
    unsigned long long v;
    f32x2() = default;
    VB200_F2 f32x2(float a) { *this = pack(a, a); }                       // broadcast
    VB200_F2 static f32x2 pack(float a, float b) {
        f32x2 r;
#if defined(__CUDA_ARCH__)
        asm("mov.b64 %0, {%1,%2};" : "=l"(r.v) : "f"(a), "f"(b));
#else
        uint32_t x, y; std::memcpy(&x, &a, 4); std::memcpy(&y, &b, 4); r.v = (unsigned long long)x | ((unsigned long long)y << 32);
#endif
        return r;
    }
    VB200_F2 float lo() const {
#if defined(__CUDA_ARCH__)
        float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a;
#else
        uint32_t x = uint32_t(v); float a; std::memcpy(&a, &x, 4); return a;
#endif
    }
    VB200_F2 float hi() const {
#if defined(__CUDA_ARCH__)
        float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b;
#else
        uint32_t x = uint32_t(v >> 32); float a; std::memcpy(&a, &x, 4); return a;
#endif
    }
};

VB200_F2 f32x2 operator+(f32x2 a, f32x2 b) {
#if defined(__CUDA_ARCH__)
    f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r;
#else
    return f32x2::pack(a.lo() + b.lo(), a.hi() + b.hi());
#endif
}
VB200_F2 f32x2 operator-(f32x2 a, f32x2 b) {
#if defined(__CUDA_ARCH__)
    f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r;
#else
    return f32x2::pack(a.lo() - b.lo(), a.hi() - b.hi());
#endif
}
VB200_F2 f32x2 operator*(f32x2 a, f32x2 b) {
#if defined(__CUDA_ARCH__)
    f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r;
#else
    return f32x2::pack(a.lo() * b.lo(), a.hi() * b.hi());
#endif
}
VB200_F2 f32x2 operator-(f32x2 a) { return f32x2::pack(-a.lo(), -a.hi()); }
VB200_F2 f32x2 operator+(f32x2 a, float b) { return a + f32x2(b); }
VB200_F2 f32x2 operator+(float a, f32x2 b) { return f32x2(a) + b; }
VB200_F2 f32x2 operator-(f32x2 a, float b) { return a - f32x2(b); }
VB200_F2 f32x2 operator-(float a, f32x2 b) { return f32x2(a) - b; }
VB200_F2 f32x2 operator*(f32x2 a, float b) { return a * f32x2(b); }
VB200_F2 f32x2 operator*(float a, f32x2 b) { return f32x2(a) * b; }
VB200_F2 f32x2& operator+=(f32x2& a, f32x2 b) { a = a + b; return a; }
VB200_F2 f32x2& operator*=(f32x2& a, f32x2 b) { a = a * b; return a; }
VB200_F2 mask2 operator<(f32x2 a, f32x2 b) { return mask2{a.lo() < b.lo(), a.hi() < b.hi()}; }
VB200_F2 mask2 operator<=(f32x2 a, f32x2 b) { return mask2{a.lo() <= b.lo(), a.hi() <= b.hi()}; }
VB200_F2 mask2 operator>(f32x2 a, f32x2 b) { return b < a; }
VB200_F2 mask2 operator>=(f32x2 a, f32x2 b) { return b <= a; }
VB200_F2 mask2 operator<(f32x2 a, float b) { return a < f32x2(b); }
VB200_F2 mask2 operator>(f32x2 a, float b) { return a > f32x2(b); }
VB200_F2 mask2 operator&&(mask2 a, mask2 b) { return mask2{a.lo && b.lo, a.hi && b.hi}; }
VB200_F2 mask2 operator||(mask2 a, mask2 b) { return mask2{a.lo || b.lo, a.hi || b.hi}; }
VB200_F2 mask2 operator!(mask2 a) { return mask2{!a.lo, !a.hi}; }

// a*b + c: fused (one FFMA2) for pairs; for floats the plain expression, so that a --fmad=false build keeps its two roundings
VB200_F2 f32x2 mad(f32x2 a, f32x2 b, f32x2 c) {
#if defined(__CUDA_ARCH__)
    f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r;
#else
    return f32x2::pack(a.lo() * b.lo() + c.lo(), a.hi() * b.hi() + c.hi());
#endif
}
VB200_F2 f32x2 mad(f32x2 a, f32x2 b, float c) { return mad(a, b, f32x2(c)); }
VB200_F2 f32x2 mad(f32x2 a, float b, f32x2 c) { return mad(a, f32x2(b), c); }
VB200_F2 f32x2 mad(float a, f32x2 b, f32x2 c) { return mad(f32x2(a), b, c); }
VB200_F2 f32x2 mad(f32x2 a, float b, float c) { return mad(a, f32x2(b), f32x2(c)); }
VB200_F2 float mad(float a, float b, float c) { return a * b + c; }
// 1 where the condition holds, 0 elsewhere
VB200_F2 f32x2 indicator(mask2 m) { return f32x2::pack(m.lo ? 1.0f : 0.0f, m.hi ? 1.0f : 0.0f); }
VB200_F2 float indicator(bool m) { return m ? 1.0f : 0.0f; }
VB200_F2 f32x2 select(mask2 m, f32x2 a, f32x2 b) { return f32x2::pack(m.lo ? a.lo() : b.lo(), m.hi ? a.hi() : b.hi()); }
VB200_F2 float select(bool m, float a, float b) { return m ? a : b; }

}} // namespace viltrum::b200
