// placeholder — replaced by the persistent greedy-heap refinement kernel
#pragma once
#include <cuda_runtime.h>
#include "../../viltrum_b200.h"
namespace viltrum { namespace b200 { namespace device {
template<class F, int DIM, bool EXACT>
inline int launch_greedy(const F&, const vb200_greedy_launch&, cudaStream_t) { return int(cudaErrorNotSupported); }
}}}
