// ldpc_toolbox_b200/csrc/layered_smem.cu — dispatch of K3q (layered_smem_impl.cuh) to its per-type
// translation units, and the shared-memory / workspace sizing the host needs.
#include "decoder_impl.hpp"

namespace ldpc {

bool launch_layered_smem_f32(const LayeredSmemLaunch& L, cudaStream_t stream);
bool launch_layered_smem_f64(const LayeredSmemLaunch& L, cudaStream_t stream);
bool launch_layered_smem_i8(const LayeredSmemLaunch& L, cudaStream_t stream);

size_t layered_smem_bytes(int n, bool is_f64, bool is_i8) {
    const size_t qsz = is_i8 ? 2 : (is_f64 ? 8 : 4);
    return (((size_t)n * qsz + 15) & ~(size_t)15) + (((size_t)n + 31) / 32) * 4 + 16;
}

size_t layered_smem_rcv_elem(bool is_f64, bool is_i8) { return is_i8 ? 1 : (is_f64 ? 8 : 4); }

bool launch_layered_smem(const LayeredSmemLaunch& L, cudaStream_t stream) {
    if (L.nframes == 0) return true;
    if (L.is_i8) return launch_layered_smem_i8(L, stream);
    return L.is_f64 ? launch_layered_smem_f64(L, stream) : launch_layered_smem_f32(L, stream);
}

}  // namespace ldpc
