// plan_tc.cuh -- tcgen05 (3xTF32) tensor-core path of the fused planner.  Placeholder until the
// kernel lands: reports "unsupported" so GMPC_PATH_AUTO always takes the FFMA path and
// GMPC_PATH_TC fails loudly.
#pragma once
#include <cuda_runtime.h>

#include "../../include/gmpc.h"
#include "common.cuh"

namespace gmpc {

struct TcState {
  bool supported = false;
  const char* why = "tensor-core kernel not built yet";
};

inline int tc_create(TcState&, const gmpc_config&, const int*, const int*, const cudaDeviceProp&) {
  return GMPC_OK;
}
inline void tc_destroy(TcState&) {}
inline int tc_set_weights(TcState&, const float* const*, const float* const*, const float* const*,
                          const float* const*, cudaStream_t, int64_t*) {
  return GMPC_OK;
}
inline bool tc_worthwhile(const TcState&, int64_t) { return false; }
inline int tc_plan(TcState&, int64_t, int, const float*, const float*, const float*, const float*,
                   int, int, float, float, float, float, float*, float*, float*, cudaStream_t,
                   int64_t*) {
  return GMPC_E_UNSUPPORTED;
}

}  // namespace gmpc
