// ldpc_toolbox_b200/csrc/bp_common.cuh — shared pieces of the frame-interleaved float / layered
// kernels (K2: flood_float_impl.cuh, K3: layered_tile_impl.cuh): 4-frame vector access to the
// [node][128] tile arrays and the early-termination bookkeeping.
#pragma once
#include <cooperative_groups.h>

#include <string>
#include <type_traits>

#include "decoder_impl.hpp"
#include "device_common.cuh"
#include "rules.cuh"

namespace ldpc {
namespace {

constexpr int kGWarps = 8;

// ---- 4-frame vector access ([node][128] arrays, lane owns frames 4*lane .. 4*lane+3) -----------
template <class T> struct V4 { T v[4]; };

template <class T> __device__ __forceinline__ V4<T> ld4(const T* base, size_t node, int lane);
template <> __device__ __forceinline__ V4<float> ld4(const float* base, size_t node, int lane) {
    float4 t = *reinterpret_cast<const float4*>(base + node * kTileFrames + lane * 4);
    return {{t.x, t.y, t.z, t.w}};
}
template <> __device__ __forceinline__ V4<double> ld4(const double* base, size_t node, int lane) {
    const double2* p = reinterpret_cast<const double2*>(base + node * kTileFrames + lane * 4);
    double2 a = p[0], b = p[1];
    return {{a.x, a.y, b.x, b.y}};
}
template <> __device__ __forceinline__ V4<int8_t> ld4(const int8_t* base, size_t node, int lane) {
    uint32_t t = *reinterpret_cast<const uint32_t*>(base + node * kTileFrames + lane * 4);
    return {{(int8_t)t, (int8_t)(t >> 8), (int8_t)(t >> 16), (int8_t)(t >> 24)}};
}
template <> __device__ __forceinline__ V4<int16_t> ld4(const int16_t* base, size_t node, int lane) {
    uint2 t = *reinterpret_cast<const uint2*>(base + node * kTileFrames + lane * 4);
    return {{(int16_t)t.x, (int16_t)(t.x >> 16), (int16_t)t.y, (int16_t)(t.y >> 16)}};
}
template <class T> __device__ __forceinline__ void st4(T* base, size_t node, int lane, const V4<T>& x);
template <> __device__ __forceinline__ void st4(float* base, size_t node, int lane, const V4<float>& x) {
    *reinterpret_cast<float4*>(base + node * kTileFrames + lane * 4) = make_float4(x.v[0], x.v[1], x.v[2], x.v[3]);
}
template <> __device__ __forceinline__ void st4(double* base, size_t node, int lane, const V4<double>& x) {
    double2* p = reinterpret_cast<double2*>(base + node * kTileFrames + lane * 4);
    p[0] = make_double2(x.v[0], x.v[1]);
    p[1] = make_double2(x.v[2], x.v[3]);
}
template <> __device__ __forceinline__ void st4(int8_t* base, size_t node, int lane, const V4<int8_t>& x) {
    uint32_t t = (uint32_t)(uint8_t)x.v[0] | (uint32_t)(uint8_t)x.v[1] << 8 | (uint32_t)(uint8_t)x.v[2] << 16 | (uint32_t)(uint8_t)x.v[3] << 24;
    *reinterpret_cast<uint32_t*>(base + node * kTileFrames + lane * 4) = t;
}
template <> __device__ __forceinline__ void st4(int16_t* base, size_t node, int lane, const V4<int16_t>& x) {
    uint2 t;
    t.x = (uint32_t)(uint16_t)x.v[0] | (uint32_t)(uint16_t)x.v[1] << 16;
    t.y = (uint32_t)(uint16_t)x.v[2] | (uint32_t)(uint16_t)x.v[3] << 16;
    *reinterpret_cast<uint2*>(base + node * kTileFrames + lane * 4) = t;
}

// ---- shared early-termination bookkeeping ------------------------------------------------------
struct StopState {
    uint32_t unsat[kLanes];
    uint32_t done[kLanes];
};


}  // namespace
}  // namespace ldpc
