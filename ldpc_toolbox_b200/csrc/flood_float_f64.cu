// ldpc_toolbox_b200/csrc/flood_float_f64.cu — one translation unit per arithmetic type so the kernels build in parallel.
#include "flood_float_impl.cuh"

namespace ldpc {
bool launch_flood_float_f64(const GenericLaunch& L, cudaStream_t s) { return launch_flood_float_t<double>(L, s); }
}  // namespace ldpc
