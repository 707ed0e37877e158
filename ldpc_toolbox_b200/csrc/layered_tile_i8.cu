// ldpc_toolbox_b200/csrc/layered_tile_i8.cu — one translation unit per arithmetic type so the kernels build in parallel.
#include "layered_tile_impl.cuh"

namespace ldpc {
bool launch_layered_tile_i8(const GenericLaunch& L, cudaStream_t s) { return launch_layered_t<float, true>(L, s); }
}  // namespace ldpc
