// Fubini adapters over the library's built-in integrands (vb200_builtin_fubini): lets C, Python (ctypes) and the tests drive the
// Fubini family — integrator_fubini<N>, regions_generator_fubini<N>, integrator_crespo2021_infinite<N> — without an nvcc TU of their
// own.  User functors get the same adapters through include/viltrum_b200/viltrum.h.
#include <viltrum_b200/device/thunks.cuh>
#include <viltrum_b200/device/fubini.cuh>
#include "builtin_integrands.cuh"
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <functional>

using namespace viltrum::b200;
using namespace viltrum::b200::builtin;

namespace {
std::mutex g_mutex;
std::unordered_map<const vb200_integrand*, std::function<void()>> g_owned;

template<class G, int N>
const vb200_integrand* own(const G& g, const char* name) {
    auto* obj = new Integrand<G, N>(g, name);
    std::lock_guard<std::mutex> lk(g_mutex);
    g_owned[obj->c_abi()] = [obj] { delete obj; };
    return obj->c_abi();
}
template<class F, int D, int N>
const vb200_integrand* finite(const char* name, const float* rmin, const float* rmax, int nrest, uint64_t m, uint64_t seed) {
    if (nrest != D - N) return nullptr;
    return own<FubiniFinite<F,N,D-N>, N>(make_fubini_finite<N, D-N>(F(), rmin, rmax, m, seed), name);
}
template<class F, int N>
const vb200_integrand* infinite(const char* name, const float* rmin, const float* rmax, int nrest, uint64_t m, uint64_t seed) {
    return own<FubiniInfinite<F,N>, N>(make_fubini_infinite<N>(F(), rmin, rmax, nrest, m, seed), name);
}
}

extern "C" const vb200_integrand* vb200_builtin_fubini(const char* name, int nfirst, const float* rest_min, const float* rest_max, int nrest,
                                                       uint64_t mc_samples, uint64_t seed) {
    if (!name || mc_samples == 0 || mc_samples > 0xffffffffull || nrest < 0 || (nrest > 0 && (!rest_min || !rest_max))) return nullptr;
    auto is = [&] (const char* n, int k) { return !std::strcmp(name, n) && nfirst == k; };
    if (is("poly3", 1))     return finite<Poly3, 3, 1>("fubini<1>(poly3)", rest_min, rest_max, nrest, mc_samples, seed);
    if (is("poly3", 2))     return finite<Poly3, 3, 2>("fubini<2>(poly3)", rest_min, rest_max, nrest, mc_samples, seed);
    if (is("shade4_16", 2)) return finite<Shade4<16>, 4, 2>("fubini<2>(shade4_16)", rest_min, rest_max, nrest, mc_samples, seed);
    if (is("shade4_64", 2)) return finite<Shade4<64>, 4, 2>("fubini<2>(shade4_64)", rest_min, rest_max, nrest, mc_samples, seed);
    if (is("shade5_16", 2)) return finite<Shade5<16>, 5, 2>("fubini<2>(shade5_16)", rest_min, rest_max, nrest, mc_samples, seed);
    if (is("shade5_16", 3)) return finite<Shade5<16>, 5, 3>("fubini<3>(shade5_16)", rest_min, rest_max, nrest, mc_samples, seed);
    if (is("decay", 1))     return infinite<Decay, 1>("fubini<1>(decay)", rest_min, rest_max, nrest, mc_samples, seed);
    if (is("decay", 2))     return infinite<Decay, 2>("fubini<2>(decay)", rest_min, rest_max, nrest, mc_samples, seed);
    if (is("walk", 1))      return infinite<WalkPlain, 1>("fubini<1>(walk)", rest_min, rest_max, nrest, mc_samples, seed);
    if (is("walk", 2))      return infinite<WalkPlain, 2>("fubini<2>(walk)", rest_min, rest_max, nrest, mc_samples, seed);
    return nullptr;
}

extern "C" void vb200_integrand_free(const vb200_integrand* f) {
    std::function<void()> del;
    {
        std::lock_guard<std::mutex> lk(g_mutex);
        auto it = g_owned.find(f);
        if (it == g_owned.end()) return;      // built-in statics and user-owned descriptors are not ours to free
        del = std::move(it->second); g_owned.erase(it);
    }
    del();
}
