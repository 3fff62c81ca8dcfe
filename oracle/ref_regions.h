// TEST INFRASTRUCTURE — not product code.
// Region-list extraction from the unmodified reference through its own Logger::log hook
// (reference src/newton-cotes/integrator-region-based.h:19 calls logger.log(seq_regions)).
#pragma once
#include "ref_common.h"

namespace vref {

template<class T, class=void> struct is_region_seq : std::false_type {};
template<class T> struct is_region_seq<T, std::void_t<decltype(std::declval<T>().begin()->range())>> : std::true_type {};

template<typename T>
struct RegionSinkT {
    T* reg_min=nullptr; T* reg_max=nullptr; T* reg_err=nullptr; uint32_t* reg_dim=nullptr; T* reg_data=nullptr;
    const void* base=nullptr;       // address of element 0 of the logged region vector
    std::size_t count=0;
};
using RegionSink = RegionSinkT<float>;

template<typename T>
class DumpLoggerT {
    RegionSinkT<T>* sink;
public:
    DumpLoggerT(RegionSinkT<T>* s) : sink(s) {}
    std::string name() const { return ""; }
    void set_name(const std::string&) {}
    template<typename Number> void log_progress(const Number&, const Number& = Number(1)) {}
    template<typename Data> void log(const Data& d) {
        if constexpr (is_region_seq<Data>::value) {
            g_phase.t_log = now_s();
            using Reg = std::decay_t<decltype(*d.begin())>;
            constexpr std::size_t D = Reg::dimensions;
            constexpr std::size_t S = Reg::rule::samples;
            std::size_t n = 0;
            std::size_t sd = 1; for (std::size_t i=0;i<D;++i) sd *= S;
            sink->base = static_cast<const void*>(&(*d.begin()));
            for (const auto& r : d) {
                for (std::size_t i=0;i<D;++i) {
                    if (sink->reg_min) sink->reg_min[n*D+i] = r.range().min(i);
                    if (sink->reg_max) sink->reg_max[n*D+i] = r.range().max(i);
                }
                if constexpr (std::is_same_v<std::decay_t<decltype(r.extra())>, std::tuple<T,std::size_t>> ||
                              std::is_same_v<std::decay_t<decltype(r.extra())>, std::tuple<double,std::size_t>>) {      // error_heuristic_mixed keys are doubles
                    if (sink->reg_err) sink->reg_err[n] = T(std::get<0>(r.extra()));
                    if (sink->reg_dim) sink->reg_dim[n] = uint32_t(std::get<1>(r.extra()));
                }
                if (sink->reg_data) {
                    // multiarray::operator[] is public; iterate in its own dim-0-fastest order (multiarray.h:20-25)
                    std::array<std::size_t,D> sres; sres.fill(S);
                    std::size_t k = 0;
                    for (auto idx : viltrum::multidimensional_range(sres)) sink->reg_data[n*sd+(k++)] = r.data[idx];
                }
                ++n;
            }
            sink->count = n;
        }
    }
};
using DumpLogger = DumpLoggerT<float>;

} // namespace vref
