// Specialised register-resident Jacobi encoder (fp32, L = 10).  Placeholder
// until the kernel lands: reports "unsupported" so the generic kernel runs.
#pragma once
#include "oc_generic.cuh"

namespace oc {
inline bool encode_fast_supported(const EncodeParams&) { return false; }
inline int encode_fast_launch(const EncodeParams&, int, cudaStream_t) { return -1; }
}  // namespace oc
