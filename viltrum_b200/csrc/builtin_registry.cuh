// Included by builtin_fast.cu (default flags) and builtin_exact.cu (--fmad=false, -DVILTRUM_B200_EXACT): defines the
// table of built-in integrands for that flavour.
#pragma once
#include <viltrum_b200/device/thunks.cuh>
#include "builtin_integrands.cuh"

namespace viltrum { namespace b200 { namespace builtin {

struct Entry { const char* name; const vb200_integrand* desc; };

#ifdef VILTRUM_B200_EXACT
#define VB200_BUILTIN_TABLE builtin_table_exact
#define VB200_BUILTIN_TABLE64 builtin_table64_exact
#else
#define VB200_BUILTIN_TABLE builtin_table_fast
#define VB200_BUILTIN_TABLE64 builtin_table64_fast
#endif

static const Entry* make_table(int* count) {
    static const Integrand<X2Y2,2> x2y2{X2Y2(), "x2y2"};
    static const Integrand<Ind2,2> ind2{Ind2(), "ind2"};
    static const Integrand<Cubic1,1> cubic1{Cubic1(), "cubic1"};
    static const Integrand<Poly3,3> poly3{Poly3(), "poly3"};
    static const Integrand<Shade4<16>,4> shade4_16{Shade4<16>(), "shade4_16"};
    static const Integrand<Shade4<64>,4> shade4_64{Shade4<64>(), "shade4_64"};
    static const Integrand<Shade5<16>,5> shade5_16{Shade5<16>(), "shade5_16"};
    static const Integrand<Shade5<64>,5> shade5_64{Shade5<64>(), "shade5_64"};
    static const Integrand<SmoothEdge2,2> smooth_edge2{SmoothEdge2(), "smooth_edge2"};
    static const InfiniteIntegrand<Walk> walk{Walk(), "walk"};
    static const InfiniteIntegrand<Decay> decay{Decay(), "decay"};
    static const InfiniteIntegrand<WalkPlain> walk_plain{WalkPlain(), "walk_plain"};
    static const InfiniteIntegrand<WalkSteps> walk_steps{WalkSteps(), "walk_steps"};
    static const InfiniteIntegrand<DecayPlain> decay_plain{DecayPlain(), "decay_plain"};
    static const Entry table[] = {
        {"x2y2", x2y2.c_abi()}, {"ind2", ind2.c_abi()}, {"cubic1", cubic1.c_abi()}, {"poly3", poly3.c_abi()},
        {"shade4_16", shade4_16.c_abi()}, {"shade4_64", shade4_64.c_abi()}, {"shade5_16", shade5_16.c_abi()},
        {"shade5_64", shade5_64.c_abi()}, {"smooth_edge2", smooth_edge2.c_abi()}, {"walk", walk.c_abi()}, {"decay", decay.c_abi()}, {"walk_plain", walk_plain.c_abi()},
        {"walk_steps", walk_steps.c_abi()}, {"decay_plain", decay_plain.c_abi()},
    };
    *count = int(sizeof(table)/sizeof(table[0]));
    return table;
}

static const Entry* make_table64(int* count) {
    static const Integrand64<X2Y2d,2> x2y2{X2Y2d(), "x2y2"};
    static const Integrand64<Ind2d,2> ind2{Ind2d(), "ind2"};
    static const Integrand64<Cubic1d,1> cubic1{Cubic1d(), "cubic1"};
    static const Integrand64<Poly3d,3> poly3{Poly3d(), "poly3"};
    static const Integrand64<SmoothEdge2d,2> smooth_edge2{SmoothEdge2d(), "smooth_edge2"};
    static const Integrand64<Shade4d<16>,4> shade4_16{Shade4d<16>(), "shade4_16"};
    static const Entry table[] = {
        {"x2y2", x2y2.c_abi()}, {"ind2", ind2.c_abi()}, {"cubic1", cubic1.c_abi()}, {"poly3", poly3.c_abi()},
        {"smooth_edge2", smooth_edge2.c_abi()}, {"shade4_16", shade4_16.c_abi()},
    };
    *count = int(sizeof(table)/sizeof(table[0]));
    return table;
}

}}}

extern "C" const viltrum::b200::builtin::Entry* VB200_BUILTIN_TABLE64(int* count) { return viltrum::b200::builtin::make_table64(count); }
extern "C" const viltrum::b200::builtin::Entry* VB200_BUILTIN_TABLE(int* count) { return viltrum::b200::builtin::make_table(count); }
