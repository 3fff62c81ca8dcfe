// Launch thunks: the only place where kernels templated on the user's functor type are instantiated.  They are
// compiled in the USER's nvcc translation unit (or, for the built-in synthetic integrands, inside the library)
// and handed to libviltrum_b200.so through the vb200_integrand table (SURVEY.md §8b "launch thunk table").
//
//   viltrum::b200::Integrand<F,DIM> g(f);          // finite: F is  float operator()(const std::array<float,DIM>&) const
//   viltrum::b200::InfiniteIntegrand<F> g(f);      // infinite: F is template<class Seq> float operator()(const Seq&) const
//   vb200_mc_per_bin(ctx, g.c_abi(), ...);
//
// Functors must be trivially copyable and __device__-callable.  Compile the TU with
//   nvcc -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a
// (add --fmad=false and define VILTRUM_B200_EXACT for the bit-exact twin).
#pragma once
#include <cstring>
#include <cstdlib>
#include <type_traits>
#include <cuda_runtime.h>
#include "../../viltrum_b200.h"
#include "mc_per_bin.cuh"
#include "mc_scatter.cuh"
#include "walk.cuh"
#include "greedy.cuh"

namespace viltrum { namespace b200 {

#ifdef VILTRUM_B200_EXACT
constexpr bool kExactTU = true;
#else
constexpr bool kExactTU = false;
#endif

namespace detail {

inline int sm_count() {
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
    return n;
}

// grid = resident CTA capacity of the chip for this kernel (SMs x occupancy), capped by the available work
template<class K>
inline int persistent_grid(K kernel, int threads, uint64_t work_items, int hint) {
    if (hint > 0) return hint;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, 0) != cudaSuccess || occ < 1) occ = 1;
    uint64_t g = uint64_t(sm_count()) * uint64_t(occ);
    if (g > work_items) g = work_items;
    if (g < 1) g = 1;
    return int(g);
}

template<class F, int DIM, bool EXACT>
struct FiniteThunks {
    static_assert(std::is_trivially_copyable<F>::value, "integrand functors cross the C ABI by value: must be trivially copyable");
    // the descriptor points at a live F owned by the Integrand object: copy-construct from it (closures and adapters are
    // not default-constructible)
    static F functor(const vb200_integrand* self) { return *static_cast<const F*>(self->functor); }

    template<int DB, bool MOMENTS, bool NARROW, int RNG>
    static int launch_mc_rng(const F& f, const vb200_mc_launch& a, cudaStream_t st) {
        auto k = device::mc_per_bin_kernel<F, DIM, DB, MOMENTS, EXACT, NARROW, RNG>;
        const uint64_t bins_per_cta = uint64_t(device::MC_THREADS) / a.lanes_per_bin;     // one warp step of every warp
        const uint64_t ctas = (a.bin_end - a.bin_begin + bins_per_cta - 1) / bins_per_cta;
        const int grid = persistent_grid(k, device::MC_THREADS, ctas, a.grid_hint);
        k<<<grid, device::MC_THREADS, 0, st>>>(f, a);
        return int(cudaGetLastError());
    }
    template<int DB, bool MOMENTS, bool NARROW>
    static int launch_mc(const F& f, const vb200_mc_launch& a, cudaStream_t st) {
        return a.rng == VB200_RNG_PHILOX ? launch_mc_rng<DB, MOMENTS, NARROW, device::MC_RNG_PHILOX>(f, a, st)
                                         : launch_mc_rng<DB, MOMENTS, NARROW, device::MC_RNG_XOSHIRO>(f, a, st);
    }
    template<int DB>
    static int mc_db(const F& f, const vb200_mc_launch& a, cudaStream_t st) {
        // a.narrow_binned: 16-bit draws for the binned dimensions (driver: every binned dimension has >= 256 bins)
        if (a.narrow_binned) return (a.sum_f || a.sum_f2) ? launch_mc<DB, true, true>(f, a, st) : launch_mc<DB, false, true>(f, a, st);
        return (a.sum_f || a.sum_f2) ? launch_mc<DB, true, false>(f, a, st) : launch_mc<DB, false, false>(f, a, st);
    }
    static int mc(const vb200_integrand* self, const void* args, void* stream) {
        const vb200_mc_launch& a = *static_cast<const vb200_mc_launch*>(args);
        const F f = functor(self); cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (a.domain.dim != DIM) return int(cudaErrorInvalidValue);
        if (a.domain.dimbins == 1) return mc_db<1>(f, a, st);
        if constexpr (DIM >= 2) { if (a.domain.dimbins == 2) return mc_db<2>(f, a, st); }
        if constexpr (DIM >= 3) { if (a.domain.dimbins == 3) return mc_db<3>(f, a, st); }
        return int(cudaErrorInvalidValue);
    }

    template<int DB>
    static int launch_replay(const F& f, const vb200_replay_launch& a, cudaStream_t st) {
        const uint64_t n = a.bin_end - a.bin_begin;
        device::mc_replay_kernel<F, DIM, DB, EXACT><<<unsigned((n + 127) / 128), 128, 0, st>>>(f, a);
        return int(cudaGetLastError());
    }
    static int replay(const vb200_integrand* self, const void* args, void* stream) {
        const vb200_replay_launch& a = *static_cast<const vb200_replay_launch*>(args);
        const F f = functor(self); cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (a.domain.dim != DIM) return int(cudaErrorInvalidValue);
        if (a.domain.dimbins == 1) return launch_replay<1>(f, a, st);
        if constexpr (DIM >= 2) { if (a.domain.dimbins == 2) return launch_replay<2>(f, a, st); }
        if constexpr (DIM >= 3) { if (a.domain.dimbins == 3) return launch_replay<3>(f, a, st); }
        return int(cudaErrorInvalidValue);
    }

    static int eval(const vb200_integrand* self, const void* args, void* stream) {
        const vb200_eval_launch& a = *static_cast<const vb200_eval_launch*>(args);
        if (a.dim != DIM || a.f64) return int(cudaErrorInvalidValue);
        if (a.n == 0) return 0;
        auto k = device::eval_points_kernel<F, DIM, float, EXACT>;
        const int grid = persistent_grid(k, 256, (a.n + 255) / 256, 0);
        k<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(functor(self), a);
        return int(cudaGetLastError());
    }

    template<int DB>
    static int launch_scatter(const F& f, const vb200_scatter_launch& a, cudaStream_t st) {
        auto k = device::mc_scatter_kernel<F, DIM, DB, EXACT>;
        const uint64_t n = (a.sample_end - a.sample_begin + 7) / 8 + 1;      // one thread per group of eight samples
        const int grid = persistent_grid(k, 256, (n + 255) / 256, a.grid_hint);
        k<<<grid, 256, 0, st>>>(f, a);
        return int(cudaGetLastError());
    }
    static int scatter(const vb200_integrand* self, const void* args, void* stream) {
        const vb200_scatter_launch& a = *static_cast<const vb200_scatter_launch*>(args);
        const F f = functor(self); cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (a.domain.dim != DIM) return int(cudaErrorInvalidValue);
        if (a.domain.dimbins == 1) return launch_scatter<1>(f, a, st);
        if constexpr (DIM >= 2) { if (a.domain.dimbins == 2) return launch_scatter<2>(f, a, st); }
        if constexpr (DIM >= 3) { if (a.domain.dimbins == 3) return launch_scatter<3>(f, a, st); }
        return int(cudaErrorInvalidValue);
    }

    static int greedy(const vb200_integrand* self, const void* args, void* stream) {
        const vb200_greedy_launch& a = *static_cast<const vb200_greedy_launch*>(args);
        if (a.dim != DIM || a.f64) return int(cudaErrorInvalidValue);
        return device::launch_greedy<F, DIM, EXACT, float>(functor(self), a, static_cast<cudaStream_t>(stream));
    }
};

// double-precision integrands (functor over std::array<double,DIM> returning double): the Newton-Cotes region family only
template<class F, int DIM, bool EXACT>
struct FiniteThunks64 {
    static_assert(std::is_trivially_copyable<F>::value, "integrand functors cross the C ABI by value: must be trivially copyable");
    static F functor(const vb200_integrand* self) { return *static_cast<const F*>(self->functor); }
    static int eval(const vb200_integrand* self, const void* args, void* stream) {
        const vb200_eval_launch& a = *static_cast<const vb200_eval_launch*>(args);
        if (a.dim != DIM || !a.f64) return int(cudaErrorInvalidValue);
        if (a.n == 0) return 0;
        auto k = device::eval_points_kernel<F, DIM, double, EXACT>;
        const int grid = persistent_grid(k, 256, (a.n + 255) / 256, 0);
        k<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(functor(self), a);
        return int(cudaGetLastError());
    }
    // Range<double,DIM> through regions-generator-adaptive-heap.h:18-45: the greedy kernel with T = double
    static int greedy(const vb200_integrand* self, const void* args, void* stream) {
        const vb200_greedy_launch& a = *static_cast<const vb200_greedy_launch*>(args);
        if (a.dim != DIM || !a.f64) return int(cudaErrorInvalidValue);
        return device::launch_greedy<F, DIM, EXACT, double>(functor(self), a, static_cast<cudaStream_t>(stream));
    }
};

template<class F, bool EXACT>
struct InfiniteThunks {
    static_assert(std::is_trivially_copyable<F>::value, "integrand functors cross the C ABI by value: must be trivially copyable");
    // the descriptor points at a live F owned by the Integrand object: copy-construct from it (closures and adapters are
    // not default-constructible)
    static F functor(const vb200_integrand* self) { return *static_cast<const F*>(self->functor); }

    template<int DB, bool MOMENTS> static auto wavefront_or_plain(std::true_type) { return device::walk_wavefront_kernel<F, DB, MOMENTS, EXACT>; }
    template<int DB, bool MOMENTS> static auto wavefront_or_plain(std::false_type) { return device::walk_kernel<F, DB, MOMENTS, EXACT>; }
    template<class K>
    static int launch_walk_kernel(K k, const F& f, const vb200_walk_launch& a, cudaStream_t st) {
        const uint64_t bins_per_cta = uint64_t(device::MC_THREADS) / a.lanes_per_bin;
        const uint64_t ctas = (a.bin_end - a.bin_begin + bins_per_cta - 1) / bins_per_cta;
        const int grid = persistent_grid(k, device::MC_THREADS, ctas, a.grid_hint);
        k<<<grid, device::MC_THREADS, 0, st>>>(f, a);
        return int(cudaGetLastError());
    }
    template<int DB, bool MOMENTS>
    static int launch_walk(const F& f, const vb200_walk_launch& a, cudaStream_t st) {
        // state machines that consume 2 + 2 elements per begin()/step() are fed whole Philox blocks (one generator call per lane and
        // iteration, immediate refill) as long as every explicit range entry sits in block 0
        if constexpr (device::has_block_steps<F>::value) {
            if (a.domain.dim <= 4 && DB <= 4) {
                // one lane per bin (every large grid): the two-tile window keeps the lanes of a warp busy across the uneven lengths of
                // their bins.  Needs enough tiles per warp to still balance the tail through the ticket counter; VB200_WALK_WINDOW=0/1
                // switches it off / forces it (tests compare both kernels bit for bit).
                if (a.lanes_per_bin == 1 && a.spp < 0xffffffffu) {
                    auto kw = a.domain.dim > DB ? device::walk_block_window_kernel<F, DB, MOMENTS, EXACT, true>
                                                : device::walk_block_window_kernel<F, DB, MOMENTS, EXACT, false>;
                    const uint64_t tiles = (a.bin_end - a.bin_begin + 31u) / 32u;
                    const uint64_t ctas = (tiles + device::MC_THREADS / 32 - 1) / (device::MC_THREADS / 32);
                    const int grid = persistent_grid(kw, device::MC_THREADS, ctas, a.grid_hint);
                    bool window = tiles >= 6ull * uint64_t(grid) * (device::MC_THREADS / 32);
                    if (const char* env = std::getenv("VB200_WALK_WINDOW")) window = env[0] != '0';
                    if (window) { kw<<<grid, device::MC_THREADS, 0, st>>>(f, a); return int(cudaGetLastError()); }
                }
                return launch_walk_kernel(device::walk_block_kernel<F, DB, MOMENTS, EXACT>, f, a, st);
            }
        }
        // functors that also describe themselves as a state machine get the wavefront kernel (lane refill)
        auto k = wavefront_or_plain<DB, MOMENTS>(std::integral_constant<bool, device::has_steps<F>::value>());
        const uint64_t bins_per_cta = uint64_t(device::MC_THREADS) / a.lanes_per_bin;
        const uint64_t ctas = (a.bin_end - a.bin_begin + bins_per_cta - 1) / bins_per_cta;
        const int grid = persistent_grid(k, device::MC_THREADS, ctas, a.grid_hint);
        k<<<grid, device::MC_THREADS, 0, st>>>(f, a);
        return int(cudaGetLastError());
    }
    template<int DB>
    static int walk_db(const F& f, const vb200_walk_launch& a, cudaStream_t st) {
        return (a.sum_f || a.sum_f2) ? launch_walk<DB, true>(f, a, st) : launch_walk<DB, false>(f, a, st);
    }
    static int walk(const vb200_integrand* self, const void* args, void* stream) {
        const vb200_walk_launch& a = *static_cast<const vb200_walk_launch*>(args);
        const F f = functor(self); cudaStream_t st = static_cast<cudaStream_t>(stream);
        switch (a.domain.dimbins) {
            case 1: return walk_db<1>(f, a, st);
            case 2: return walk_db<2>(f, a, st);
            case 3: return walk_db<3>(f, a, st);
        }
        return int(cudaErrorInvalidValue);
    }
    static int walk_replay(const vb200_integrand* self, const void* args, void* stream) {
        const vb200_walk_replay_launch& a = *static_cast<const vb200_walk_replay_launch*>(args);
        const uint64_t n = a.bin_end - a.bin_begin;
        const unsigned grid = unsigned((n + 127) / 128); cudaStream_t st = static_cast<cudaStream_t>(stream);
        switch (a.domain.dimbins) {
            case 1: device::walk_replay_kernel<F, 1, EXACT><<<grid, 128, 0, st>>>(functor(self), a); break;
            case 2: device::walk_replay_kernel<F, 2, EXACT><<<grid, 128, 0, st>>>(functor(self), a); break;
            case 3: device::walk_replay_kernel<F, 3, EXACT><<<grid, 128, 0, st>>>(functor(self), a); break;
            default: return int(cudaErrorInvalidValue);
        }
        return int(cudaGetLastError());
    }
    template<int DB>
    static int launch_walk_scatter(const F& f, const vb200_scatter_launch& a, cudaStream_t st) {
        auto k = device::walk_scatter_kernel<F, DB, EXACT>;
        const uint64_t n = a.sample_end - a.sample_begin;
        const int grid = persistent_grid(k, 256, (n + 255) / 256, a.grid_hint);
        k<<<grid, 256, 0, st>>>(f, a);
        return int(cudaGetLastError());
    }
    static int walk_scatter(const vb200_integrand* self, const void* args, void* stream) {
        const vb200_scatter_launch& a = *static_cast<const vb200_scatter_launch*>(args);
        const F f = functor(self); cudaStream_t st = static_cast<cudaStream_t>(stream);
        switch (a.domain.dimbins) {
            case 1: return launch_walk_scatter<1>(f, a, st);
            case 2: return launch_walk_scatter<2>(f, a, st);
            case 3: return launch_walk_scatter<3>(f, a, st);
        }
        return int(cudaErrorInvalidValue);
    }
};

} // namespace detail

// Owns a copy of the functor and the C-ABI descriptor that points at it.
template<class F, int DIM, bool EXACT = kExactTU>
class Integrand {
    F f_; vb200_integrand d_;
    void bind(const char* name) {
        std::memset(&d_, 0, sizeof(d_));
        d_.abi_version = VB200_ABI_VERSION; d_.dim = DIM; d_.functor = &f_; d_.functor_bytes = uint32_t(sizeof(F));
        d_.flags = EXACT ? VB200_INTEGRAND_EXACT : 0u; d_.name = name;
        using T = detail::FiniteThunks<F, DIM, EXACT>;
        d_.launch[VB200_K_MC_PER_BIN] = &T::mc;
        d_.launch[VB200_K_MC_REPLAY] = &T::replay;
        d_.launch[VB200_K_EVAL_POINTS] = &T::eval;
        d_.launch[VB200_K_MC_SCATTER] = &T::scatter;
        d_.launch[VB200_K_ADAPTIVE_EXACT] = &T::greedy;
    }
public:
    explicit Integrand(const F& f, const char* name = "user integrand") : f_(f) { bind(name); }
    Integrand(const Integrand& o) : f_(o.f_) { bind(o.d_.name); }
    Integrand& operator=(const Integrand& o) { f_ = o.f_; bind(o.d_.name); return *this; }
    const vb200_integrand* c_abi() const { return &d_; }
};

// double-precision integrand: double operator()(const std::array<double,DIM>&) const
template<class F, int DIM, bool EXACT = kExactTU>
class Integrand64 {
    F f_; vb200_integrand d_;
    void bind(const char* name) {
        std::memset(&d_, 0, sizeof(d_));
        d_.abi_version = VB200_ABI_VERSION; d_.dim = DIM; d_.functor = &f_; d_.functor_bytes = uint32_t(sizeof(F));
        d_.flags = (EXACT ? VB200_INTEGRAND_EXACT : 0u) | VB200_INTEGRAND_F64; d_.name = name;
        d_.launch[VB200_K_EVAL_POINTS] = &detail::FiniteThunks64<F, DIM, EXACT>::eval;
        d_.launch[VB200_K_ADAPTIVE_EXACT] = &detail::FiniteThunks64<F, DIM, EXACT>::greedy;
    }
public:
    explicit Integrand64(const F& f, const char* name = "user integrand (f64)") : f_(f) { bind(name); }
    Integrand64(const Integrand64& o) : f_(o.f_) { bind(o.d_.name); }
    Integrand64& operator=(const Integrand64& o) { f_ = o.f_; bind(o.d_.name); return *this; }
    const vb200_integrand* c_abi() const { return &d_; }
};

template<class F, bool EXACT = kExactTU>
class InfiniteIntegrand {
    F f_; vb200_integrand d_;
    void bind(const char* name) {
        std::memset(&d_, 0, sizeof(d_));
        d_.abi_version = VB200_ABI_VERSION; d_.dim = -1; d_.functor = &f_; d_.functor_bytes = uint32_t(sizeof(F));
        d_.flags = EXACT ? VB200_INTEGRAND_EXACT : 0u; d_.name = name;
        using T = detail::InfiniteThunks<F, EXACT>;
        d_.launch[VB200_K_WALK] = &T::walk;
        d_.launch[VB200_K_WALK_REPLAY] = &T::walk_replay;
        d_.launch[VB200_K_WALK_SCATTER] = &T::walk_scatter;
    }
public:
    explicit InfiniteIntegrand(const F& f, const char* name = "user integrand") : f_(f) { bind(name); }
    InfiniteIntegrand(const InfiniteIntegrand& o) : f_(o.f_) { bind(o.d_.name); }
    InfiniteIntegrand& operator=(const InfiniteIntegrand& o) { f_ = o.f_; bind(o.d_.name); return *this; }
    const vb200_integrand* c_abi() const { return &d_; }
};

}} // namespace viltrum::b200
