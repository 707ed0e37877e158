// ldpc_toolbox_b200/csrc/layered_smem_f32.cu — one translation unit per arithmetic type so the kernels build in parallel.
#include "layered_smem_impl.cuh"

namespace ldpc {
bool launch_layered_smem_f32(const LayeredSmemLaunch& L, cudaStream_t stream) {
    switch (L.rule) {
        case kPhi: return launch_t<float, kPhi, false, false>(L, stream);
        case kTanh: return launch_t<float, kTanh, false, false>(L, stream);
        case kMinstarapprox: return launch_t<float, kMinstarapprox, false, false>(L, stream);
        case kMinstarapproxExact: return launch_t<float, kMinstarapproxExact, false, false>(L, stream);
        case kAminstarExact: return launch_t<float, kAminstarExact, false, false>(L, stream);
        default: return launch_t<float, kAminstar, false, false>(L, stream);
    }
}
}  // namespace ldpc
