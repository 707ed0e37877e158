// ldpc_toolbox_b200/csrc/layered_smem_i8.cu — one translation unit per arithmetic type so the kernels build in parallel.
#include "layered_smem_impl.cuh"

namespace ldpc {
bool launch_layered_smem_i8(const LayeredSmemLaunch& L, cudaStream_t stream) {
    if (L.rule == kMinstarapprox) return L.hardlimit ? launch_t<float, kMinstarapprox, true, true>(L, stream) : launch_t<float, kMinstarapprox, true, false>(L, stream);
    return L.hardlimit ? launch_t<float, kAminstar, true, true>(L, stream) : launch_t<float, kAminstar, true, false>(L, stream);
}
}  // namespace ldpc
