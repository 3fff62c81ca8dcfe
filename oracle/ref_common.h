// TEST INFRASTRUCTURE — not product code.
//
// Common plumbing for the harness that drives the UNMODIFIED reference (headers included from
// /root/reference where they lie; see oracle/Makefile).  Nothing from the reference is copied here.
// `#define private public` is the classic white-box test trick: Region::data has no accessor
// (reference src/newton-cotes/region.h:36) and the parity tests need the stored samples bit for bit.
#pragma once
#include <array>
#include <vector>
#include <tuple>
#include <string>
#include <cstring>
#include <cstdint>
#include <cmath>
#include <random>
#include <thread>
#include <mutex>
#include <iostream>
#include <iomanip>
#include <chrono>
#include <algorithm>
#include <numeric>
#include <execution>
#include <list>
#include <complex>
#include <limits>
#include <type_traits>
#include <cassert>
#include <stdexcept>
#include <functional>

#define private public
#include "viltrum.h"      // -I/root/reference
#undef private

#include "integrands.h"
#include "oracle_api.h"

namespace vref {

// Phase clock of the timing legs (bench.py cpu_baseline / --impl reference): every region-based entry point notes when it
// started, when the reference handed its region list to Logger::log (= generation done, integrator-region-based.h:19) and when
// it returned; vo_phase_times() reads the three back.
struct PhaseClock { double t_begin = 0, t_log = 0, t_end = 0; };
inline PhaseClock g_phase;
inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct PhaseTimer {
    PhaseTimer() { g_phase.t_begin = now_s(); g_phase.t_log = g_phase.t_end = 0; }
    ~PhaseTimer() { g_phase.t_end = now_s(); if (g_phase.t_log == 0) g_phase.t_log = g_phase.t_begin; }
};

template<std::size_t DB>
inline std::size_t tensor_pos(const std::array<std::size_t,DB>& p, const std::array<std::size_t,DB>& res) {
    std::size_t pos=0, prod=1;
    for (std::size_t d=0; d<DB; ++d) { pos += p[d]*prod; prod *= res[d]; }
    return pos;
}

// Order in which reference for_each(parallel, multidimensional_range(res), f) visits the bins when the
// PSTL backend is serial: dim 0 is the task index (outer), the remaining dims iterate inside each task
// with the lowest remaining dim fastest (reference src/foreach.h:56-67).
template<std::size_t DB>
inline std::vector<std::size_t> parallel_visit_order(const std::array<std::size_t,DB>& res) {
    std::vector<std::size_t> order;
    if constexpr (DB == 1) {
        for (std::size_t d=0; d<res[0]; ++d) order.push_back(d);
    } else {
        std::array<std::size_t,DB-1> b;
        for (std::size_t s=1;s<DB;++s) b[s-1]=res[s];
        for (std::size_t d=0; d<res[0]; ++d)
            for (auto pos : viltrum::multidimensional_range(b)) {
                std::array<std::size_t,DB> full; full[0]=d;
                for (std::size_t s=1;s<DB;++s) full[s]=pos[s-1];
                order.push_back(tensor_pos(full,res));
            }
    }
    return order;
}

template<std::size_t DB>
inline std::array<std::size_t,DB> res_array(const uint64_t* res) {
    std::array<std::size_t,DB> r; for (std::size_t i=0;i<DB;++i) r[i]=std::size_t(res[i]); return r;
}
template<std::size_t D>
inline viltrum::Range<float,D> range_array(const float* rmin, const float* rmax) {
    std::array<float,D> a,b; for (std::size_t i=0;i<D;++i) { a[i]=rmin[i]; b[i]=rmax[i]; }
    return viltrum::range(a,b);
}

// Calls fn(F{}) with the integrand functor named `name`; returns -1 if unknown / not finite-dimensional.
template<typename Fn>
inline int dispatch_finite(const char* name, Fn&& fn) {
    if (!std::strcmp(name,"x2y2"))         return fn(vo::X2Y2());
    if (!std::strcmp(name,"ind2"))         return fn(vo::Ind2());
    if (!std::strcmp(name,"cubic1"))       return fn(vo::Cubic1());
    if (!std::strcmp(name,"poly3"))        return fn(vo::Poly3());
    if (!std::strcmp(name,"shade4_64"))    return fn(vo::Shade4<64>());
    if (!std::strcmp(name,"shade4_16"))    return fn(vo::Shade4<16>());
    if (!std::strcmp(name,"shade5_64"))    return fn(vo::Shade5<64>());
    if (!std::strcmp(name,"shade5_16"))    return fn(vo::Shade5<16>());
    if (!std::strcmp(name,"smooth_edge2")) return fn(vo::SmoothEdge2());
    return -1;
}

// Calls fn(F{}, integral_constant<DB>{}) for DB in {1,2} (DB <= dim).
template<typename Fn>
inline int dispatch_finite_bins(const char* name, int dimbins, Fn&& fn) {
    return dispatch_finite(name, [&] (auto f) -> int {
        using F = decltype(f);
        if (dimbins == 1) return fn(f, std::integral_constant<std::size_t,1>());
        if constexpr (F::dim >= 2) {
            if (dimbins == 2) return fn(f, std::integral_constant<std::size_t,2>());
        }
        return -2;
    });
}

// double family (Range<double,DIM>)
template<typename Fn>
inline int dispatch_finite_d(const char* name, int dimbins, Fn&& fn) {
    auto go = [&] (auto f) -> int {
        using F = decltype(f);
        if (dimbins == 1) return fn(f, std::integral_constant<std::size_t,1>());
        if constexpr (F::dim >= 2) { if (dimbins == 2) return fn(f, std::integral_constant<std::size_t,2>()); }
        return -2;
    };
    if (!std::strcmp(name,"x2y2"))         return go(vo::X2Y2T<double>());
    if (!std::strcmp(name,"ind2"))         return go(vo::Ind2T<double>());
    if (!std::strcmp(name,"cubic1"))       return go(vo::Cubic1T<double>());
    if (!std::strcmp(name,"poly3"))        return go(vo::Poly3T<double>());
    if (!std::strcmp(name,"smooth_edge2")) return go(vo::SmoothEdge2T<double>());
    if (!std::strcmp(name,"shade4_16"))    return go(vo::Shade4T<double,16>());
    return -1;
}

template<typename Fn>
inline int dispatch_infinite(const char* name, Fn&& fn) {
    if (!std::strcmp(name,"walk"))  return fn(vo::Walk());
    if (!std::strcmp(name,"decay")) return fn(vo::Decay());
    return -1;
}

} // namespace vref
