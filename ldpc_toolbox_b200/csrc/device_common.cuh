// ldpc_toolbox_b200/csrc/device_common.cuh — shared device-side definitions.
//
// HBM data layout ("frame-interleaved tiles"): all frames share one Tanner graph,
// so the frame index is the fastest-varying dimension.  A tile is 128 frames = 32 lanes x 4
// frames; every per-edge / per-variable quantity of a tile is one 128-byte line
// [node][lane] (uint32 = the 4 int8 values of the lane's 4 frames) and every gather by edge or
// variable index — identical across frames — is one fully coalesced line per warp.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ldpc {

constexpr int kTileFrames = 128;     // frames per tile (32 lanes x 4 bytes)
constexpr int kLanes = 32;

struct DeviceGraph {
    int n, m, E;
    const int* row_ptr;    // m+1
    const int* col_idx;    // E   variable of each row-major edge
    const int* col_ptr;    // n+1
    const int* col_edge;   // E   row-major edge id of each column-major slot (cols[v] order)
};

__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p) { return __ldcg(p); }
__device__ __forceinline__ void st_stream(uint32_t* p, uint32_t v) { __stcg(p, v); }

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }

// 4 bytes of 0/1 (one per frame) -> 4-bit mask
__device__ __forceinline__ uint32_t pack_bits4(uint32_t bytes01) { return (bytes01 * 0x01020408u) >> 24; }

// input_llr_quantize of the int8 arithmetics (reference src/decoder/arithmetic.rs:690-699): x8, saturate
// at +-127, round half away from zero, NaN -> 0
__device__ __forceinline__ int quantize_i8(double llr) {
    double x = 8.0 * llr;
    if (x >= 127.0) return 127;
    if (x <= -127.0) return -127;
    if (x != x) return 0;              // Rust `as i8` maps NaN to 0
    return (int)round(x);              // f64::round: half away from zero
}
__device__ __forceinline__ int quantize_i8(float llr) {
    float x = 8.0f * llr;              // exact scaling: same value as the reference's f64 product
    if (x >= 127.0f) return 127;
    if (x <= -127.0f) return -127;
    if (x != x) return 0;
    return (int)roundf(x);
}

#define LDPC_CUDA_CHECK(expr)                                                                   \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            ::ldpc::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e));         \
            return false;                                                                       \
        }                                                                                       \
    } while (0)

}  // namespace ldpc
