// ldpc_toolbox_b200/csrc/layered_tile_f32.cu — one translation unit per arithmetic type so the kernels build in parallel.
#include "layered_tile_impl.cuh"

namespace ldpc {
bool launch_layered_tile_f32(const GenericLaunch& L, cudaStream_t s) { return launch_layered_t<float, false>(L, s); }
}  // namespace ldpc
