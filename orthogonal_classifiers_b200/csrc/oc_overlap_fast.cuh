// Specialised overlap kernel (fp32, classifier resident in shared memory).
// Placeholder until the kernel lands.
#pragma once
#include "oc_generic.cuh"

namespace oc {
inline bool overlap_fast_supported(const OverlapParams&) { return false; }
inline int overlap_fast_launch(const OverlapParams&, int, cudaStream_t) { return -1; }
}  // namespace oc
