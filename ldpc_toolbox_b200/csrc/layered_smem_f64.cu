// ldpc_toolbox_b200/csrc/layered_smem_f64.cu — one translation unit per arithmetic type so the kernels build in parallel.
#include "layered_smem_impl.cuh"

namespace ldpc {
bool launch_layered_smem_f64(const LayeredSmemLaunch& L, cudaStream_t stream) {
    switch (L.rule) {
        case kPhi: return launch_t<double, kPhi, false, false>(L, stream);
        case kTanh: return launch_t<double, kTanh, false, false>(L, stream);
        case kMinstarapprox: case kMinstarapproxExact: return launch_t<double, kMinstarapprox, false, false>(L, stream);
        default: return launch_t<double, kAminstar, false, false>(L, stream);
    }
}
}  // namespace ldpc
