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
#include <type_traits>

#include "decoder_impl.hpp"
#include "device_common.cuh"

namespace ldpc {
namespace {

constexpr int kChunk = 32;       // variables per CTA
constexpr int kIngestWarps = 8;

// NW = words per lane of the decoder tile: a tile holds TF = 128*NW frames, stored per node as TF
// consecutive values (frame index = lane*4*NW + word*4 + byte).
template <typename TIn, int MODE, int NW>   // MODE 0: int8, 1: f32, 2: f64, 3: int16 decoder state
__global__ void __launch_bounds__(kIngestWarps * 32) ingest_kernel(IngestLaunch p) {
    constexpr int TF = kTileFrames * NW, ST = TF + 4;
    __shared__ __align__(16) uint8_t s_q[MODE == 0 ? kChunk * ST : 4];
    __shared__ __align__(16) uint8_t s_raw[kChunk * ST];
    __shared__ float s_f[MODE == 1 ? kChunk * (TF + 1) : 1];
    __shared__ int16_t s_h[MODE == 3 ? kChunk * (TF + 2) : 2];
    __shared__ double s_d[MODE == 2 ? kChunk * (TF + 1) : 1];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t tile = blockIdx.x;
    const int v0 = blockIdx.y * kChunk, v = v0 + lane;
    const TIn* llrs = static_cast<const TIn*>(p.llrs);
    int src = -1;
    if (v < p.n) src = p.src_map ? __ldg(p.src_map + v) : v;

    for (int fr = warp; fr < TF; fr += kIngestWarps) {
        size_t frame = tile * TF + fr;
        TIn x = TIn(1);                                  // padding frames: clean all-zero codeword
        if (frame < p.nframes) x = src >= 0 ? llrs[frame * p.llrs_len + (size_t)src] : TIn(0);
        s_raw[lane * ST + fr] = x <= TIn(0) ? 1 : 0;
        if (MODE == 0) s_q[lane * ST + fr] = (uint8_t)(quantize_i8(x) + 128);      // offset binary (flood_i8.cu)
        if (MODE == 1) s_f[lane * (TF + 1) + fr] = (float)x;
        if (MODE == 2) s_d[lane * (TF + 1) + fr] = (double)x;
        if (MODE == 3) s_h[lane * (TF + 2) + fr] = (int16_t)quantize_i8(x);
    }
    __syncthreads();
    for (int vv = warp; vv < kChunk; vv += kIngestWarps) {
        if (v0 + vv >= p.n) break;
        size_t node = tile * (size_t)p.n + (size_t)(v0 + vv);
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < NW; ++i)
                p.inq_i8[node * (TF / 4) + i * 32 + lane] = *reinterpret_cast<const uint32_t*>(&s_q[vv * ST + (i * 32 + lane) * 4]);
        }
        if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < TF / 32; ++i) p.in_f32[node * TF + i * 32 + lane] = s_f[vv * (TF + 1) + i * 32 + lane];
        }
        if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < TF / 32; ++i) p.in_f64[node * TF + i * 32 + lane] = s_d[vv * (TF + 1) + i * 32 + lane];
        }
        if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < TF / 32; ++i) p.in_i16[node * TF + i * 32 + lane] = s_h[vv * (TF + 2) + i * 32 + lane];
        }
        uint32_t hb = 0;
#pragma unroll
        for (int q = 0; q < NW; ++q)
            hb |= pack_bits4(*reinterpret_cast<const uint32_t*>(&s_raw[vv * ST + (lane * NW + q) * 4])) << (4 * q);
        if (NW == 1) static_cast<uint8_t*>(p.hard)[node * kLanes + lane] = (uint8_t)hb;
        else static_cast<uint16_t*>(p.hard)[node * kLanes + lane] = (uint16_t)hb;
    }
}

constexpr int kEmitVars = 128;     // variables per CTA: every frame row is written 128 bytes at a time

// A CTA turns the decision masks of 128 variables of one tile into 128 consecutive output bytes of each of the tile's
// frames: masks are staged transposed in shared memory ([lane][variable], so the four masks a thread needs are adjacent),
// then a thread writes four 0/1 bytes of one frame as one 32-bit store (a warp = one full 128-byte row segment).
template <int NW>
__global__ void __launch_bounds__(256) emit_kernel(EmitLaunch p) {
    using HB = typename std::conditional<NW == 1, uint8_t, uint16_t>::type;
    constexpr int TF = kTileFrames * NW, FPL = 4 * NW;              // frames per tile / per lane
    __shared__ HB s_m[kLanes][kEmitVars + 4];
    const size_t tile = blockIdx.x;
    const size_t v0 = (size_t)blockIdx.y * kEmitVars;
    const HB* fin = static_cast<const HB*>(p.final_hard) + tile * (size_t)p.n * kLanes;
    for (int idx = threadIdx.x; idx < kEmitVars * kLanes; idx += blockDim.x) {
        const int v = idx >> 5, ln = idx & 31;
        s_m[ln][v] = v0 + v < p.out_len ? fin[(v0 + v) * kLanes + ln] : (HB)0;
    }
    __syncthreads();
    const bool aligned = (((uintptr_t)p.out | p.out_stride | v0) & 3u) == 0;
    for (int idx = threadIdx.x; idx < TF * (kEmitVars / 4); idx += blockDim.x) {
        const int fr = idx >> 5, g4 = idx & 31;                       // frame of the tile, group of four variables
        const size_t frame = tile * TF + fr;
        if (frame >= p.nframes) break;
        const int ln = fr / FPL, bit = fr % FPL;
        const size_t v = v0 + 4 * (size_t)g4;
        if (v >= p.out_len) continue;
        uint32_t w = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) w |= (uint32_t)((s_m[ln][4 * g4 + i] >> bit) & 1u) << (8 * i);
        uint8_t* dst = p.out + frame * p.out_stride + v;
        if (aligned && v + 4 <= p.out_len) *reinterpret_cast<uint32_t*>(dst) = w;
        else
            for (int i = 0; i < 4 && v + i < p.out_len; ++i) dst[i] = (uint8_t)(w >> (8 * i));
    }
}

template <class F>
__global__ void emit_posteriors_kernel(const F* __restrict__ post, int n, size_t nframes, double* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;          // one thread per (frame, variable)
    if (i >= nframes * (size_t)n) return;
    const size_t frame = i / (size_t)n, v = i % (size_t)n;
    out[i] = (double)post[((frame / kTileFrames) * (size_t)n + v) * kTileFrames + frame % kTileFrames];
}

}  // namespace

bool launch_emit_posteriors(const void* post_tiles, bool is_f64, int n, size_t nframes, double* out, cudaStream_t stream) {
    const size_t total = nframes * (size_t)n;
    if (total == 0) return true;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (is_f64) emit_posteriors_kernel<double><<<blocks, 256, 0, stream>>>(static_cast<const double*>(post_tiles), n, nframes, out);
    else emit_posteriors_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const float*>(post_tiles), n, nframes, out);
    LDPC_CUDA_CHECK(cudaGetLastError());
    return true;
}

bool launch_ingest(const IngestLaunch& L, cudaStream_t stream) {
    if (L.num_tiles == 0 || L.n == 0) return true;
    dim3 grid((unsigned)L.num_tiles, (unsigned)((L.n + kChunk - 1) / kChunk)), block(kIngestWarps * 32);
    const int mode = L.inq_i8 ? 0 : (L.in_f32 ? 1 : (L.in_f64 ? 2 : 3));
    if (L.words_per_lane == 4 && mode != 0) { set_last_error("512-frame tiles are only used by the int8 decoders"); return false; }
    if (L.is_f64) {
        if (mode == 0 && L.words_per_lane == 4) ingest_kernel<double, 0, 4><<<grid, block, 0, stream>>>(L);
        else if (mode == 0) ingest_kernel<double, 0, 1><<<grid, block, 0, stream>>>(L);
        else if (mode == 1) ingest_kernel<double, 1, 1><<<grid, block, 0, stream>>>(L);
        else if (mode == 2) ingest_kernel<double, 2, 1><<<grid, block, 0, stream>>>(L);
        else ingest_kernel<double, 3, 1><<<grid, block, 0, stream>>>(L);
    } else {
        if (mode == 0 && L.words_per_lane == 4) ingest_kernel<float, 0, 4><<<grid, block, 0, stream>>>(L);
        else if (mode == 0) ingest_kernel<float, 0, 1><<<grid, block, 0, stream>>>(L);
        else if (mode == 1) ingest_kernel<float, 1, 1><<<grid, block, 0, stream>>>(L);
        else if (mode == 2) ingest_kernel<float, 2, 1><<<grid, block, 0, stream>>>(L);
        else ingest_kernel<float, 3, 1><<<grid, block, 0, stream>>>(L);
    }
    LDPC_CUDA_CHECK(cudaGetLastError());
    return true;
}

bool launch_emit(const EmitLaunch& L, cudaStream_t stream) {
    if (L.num_tiles == 0 || L.out_len == 0) return true;
    dim3 grid((unsigned)L.num_tiles, (unsigned)((L.out_len + kEmitVars - 1) / kEmitVars)), block(256);
    if (L.words_per_lane == 4) emit_kernel<4><<<grid, block, 0, stream>>>(L);
    else emit_kernel<1><<<grid, block, 0, stream>>>(L);
    LDPC_CUDA_CHECK(cudaGetLastError());
    return true;
}

}  // namespace ldpc
