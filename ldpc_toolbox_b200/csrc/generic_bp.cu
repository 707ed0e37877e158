// ldpc_toolbox_b200/csrc/generic_bp.cu — dispatch of K2 (flooding, float rules) and K3 (horizontal layered,
// frame-interleaved tiles) to their per-type translation units.
#include "decoder_impl.hpp"
#include "rules.cuh"

namespace ldpc {

bool launch_flood_float_f32(const GenericLaunch& L, cudaStream_t s);
bool launch_flood_float_f64(const GenericLaunch& L, cudaStream_t s);
bool launch_layered_tile_f32(const GenericLaunch& L, cudaStream_t s);
bool launch_layered_tile_f64(const GenericLaunch& L, cudaStream_t s);
bool launch_layered_tile_i8(const GenericLaunch& L, cudaStream_t s);

bool launch_flood_float(const GenericLaunch& L, cudaStream_t stream) {
    return L.is_f64 ? launch_flood_float_f64(L, stream) : launch_flood_float_f32(L, stream);
}

bool launch_layered(const GenericLaunch& L, cudaStream_t stream) {
    if (L.is_i8) return launch_layered_tile_i8(L, stream);
    return L.is_f64 ? launch_layered_tile_f64(L, stream) : launch_layered_tile_f32(L, stream);
}

int generic_max_row_degree() { return kRuleMaxD; }

}  // namespace ldpc
