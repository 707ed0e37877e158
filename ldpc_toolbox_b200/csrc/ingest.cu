// ldpc_toolbox_b200/csrc/ingest.cu — the two format changes at the edge of the hot path.
//
// ingest: caller layout [frame][llr] (f32/f64, what decode(llrs, ..) receives, reference
//   src/decoder.rs:29-34) -> frame-interleaved tiles, fused with
//     * depuncturing (zeros for punctured blocks, reference src/simulation/puncturing.rs:85-101),
//     * input_llr_quantize (reference src/decoder/arithmetic.rs:690-699; `as f32` :194-196),
//     * the raw-sign hard decision x <= 0.0 used by the pre-check (reference flooding.rs:57).
// emit:   final hard-decision plane -> [frame][bit] one byte per bit (DecoderOutput.codeword,
//   reference src/decoder.rs:39-48), first out_len bits of each frame (c_api/decoder.rs:61).
#include <string>

#include "decoder_impl.hpp"
#include "device_common.cuh"

namespace ldpc {
namespace {

constexpr int kChunk = 32;       // variables per CTA
constexpr int kIngestWarps = 8;

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

template <typename TIn, int MODE>   // MODE 0: int8, 1: f32, 2: f64 decoder state
__global__ void __launch_bounds__(kIngestWarps * 32) ingest_kernel(IngestLaunch p) {
    __shared__ __align__(16) uint8_t s_q[MODE == 0 ? kChunk * 132 : 4];
    __shared__ __align__(16) uint8_t s_raw[kChunk * 132];
    __shared__ float s_f[MODE == 1 ? kChunk * 129 : 1];
    __shared__ double s_d[MODE == 2 ? kChunk * 129 : 1];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t tile = blockIdx.x;
    const int v0 = blockIdx.y * kChunk, v = v0 + lane;
    const TIn* llrs = static_cast<const TIn*>(p.llrs);
    int src = -1;
    if (v < p.n) src = p.src_map ? __ldg(p.src_map + v) : v;

    for (int fr = warp; fr < kTileFrames; fr += kIngestWarps) {
        size_t frame = tile * kTileFrames + fr;
        TIn x = TIn(1);                                  // padding frames: clean all-zero codeword
        if (frame < p.nframes) x = src >= 0 ? llrs[frame * p.llrs_len + (size_t)src] : TIn(0);
        s_raw[lane * 132 + fr] = x <= TIn(0) ? 1 : 0;
        if (MODE == 0) s_q[lane * 132 + fr] = (uint8_t)(int8_t)quantize_i8(x);
        if (MODE == 1) s_f[lane * 129 + fr] = (float)x;
        if (MODE == 2) s_d[lane * 129 + fr] = (double)x;
    }
    __syncthreads();
    for (int vv = warp; vv < kChunk; vv += kIngestWarps) {
        if (v0 + vv >= p.n) break;
        size_t node = tile * (size_t)p.n + (size_t)(v0 + vv);
        if (MODE == 0) p.inq_i8[node * kLanes + lane] = *reinterpret_cast<const uint32_t*>(&s_q[vv * 132 + lane * 4]);
        if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) p.in_f32[node * kTileFrames + i * 32 + lane] = s_f[vv * 129 + i * 32 + lane];
        }
        if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) p.in_f64[node * kTileFrames + i * 32 + lane] = s_d[vv * 129 + i * 32 + lane];
        }
        p.hard[node * kLanes + lane] = (uint8_t)pack_bits4(*reinterpret_cast<const uint32_t*>(&s_raw[vv * 132 + lane * 4]));
    }
}

constexpr int kEmitChunk = 128;

__global__ void __launch_bounds__(256) emit_kernel(EmitLaunch p) {
    __shared__ uint8_t s[kTileFrames * 132];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t tile = blockIdx.x;
    const size_t v0 = (size_t)blockIdx.y * kEmitChunk;
    for (int vv = warp; vv < kEmitChunk; vv += 8) {
        size_t v = v0 + vv;
        uint32_t b = 0;
        if (v < p.out_len) b = p.final_hard[(tile * (size_t)p.n + v) * kLanes + lane];
#pragma unroll
        for (int k = 0; k < 4; ++k) s[(lane * 4 + k) * 132 + vv] = (b >> k) & 1;
    }
    __syncthreads();
    for (int fr = warp; fr < kTileFrames; fr += 8) {
        size_t frame = tile * kTileFrames + fr;
        if (frame >= p.nframes) break;
        uint8_t* o = p.out + frame * p.out_stride;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            size_t v = v0 + i * 32 + lane;
            if (v < p.out_len) o[v] = s[fr * 132 + i * 32 + lane];
        }
    }
}

}  // namespace

bool launch_ingest(const IngestLaunch& L, cudaStream_t stream) {
    if (L.num_tiles == 0 || L.n == 0) return true;
    dim3 grid((unsigned)L.num_tiles, (unsigned)((L.n + kChunk - 1) / kChunk)), block(kIngestWarps * 32);
    const int mode = L.inq_i8 ? 0 : (L.in_f32 ? 1 : 2);
    if (L.is_f64) {
        if (mode == 0) ingest_kernel<double, 0><<<grid, block, 0, stream>>>(L);
        else if (mode == 1) ingest_kernel<double, 1><<<grid, block, 0, stream>>>(L);
        else ingest_kernel<double, 2><<<grid, block, 0, stream>>>(L);
    } else {
        if (mode == 0) ingest_kernel<float, 0><<<grid, block, 0, stream>>>(L);
        else if (mode == 1) ingest_kernel<float, 1><<<grid, block, 0, stream>>>(L);
        else ingest_kernel<float, 2><<<grid, block, 0, stream>>>(L);
    }
    LDPC_CUDA_CHECK(cudaGetLastError());
    return true;
}

bool launch_emit(const EmitLaunch& L, cudaStream_t stream) {
    if (L.num_tiles == 0 || L.out_len == 0) return true;
    dim3 grid((unsigned)L.num_tiles, (unsigned)((L.out_len + kEmitChunk - 1) / kEmitChunk)), block(256);
    emit_kernel<<<grid, block, 0, stream>>>(L);
    LDPC_CUDA_CHECK(cudaGetLastError());
    return true;
}

}  // namespace ldpc
