// ldpc_toolbox_b200/csrc/decoder.hpp — C++ mirror of the reference's decoder plugin boundary.
//
//   trait LdpcDecoder { fn decode(&mut self, llrs: &[f64], max_iterations: usize)
//                         -> Result<DecoderOutput, DecoderOutput> }      reference src/decoder.rs:19-35
//   struct DecoderOutput { codeword: Vec<u8>, iterations: usize }        reference src/decoder.rs:39-48
//   trait DecoderFactory { fn build_decoder(&self, h: SparseMatrix) -> Box<dyn LdpcDecoder> }
//                                                                         reference src/decoder/factory.rs:19-25
//
// The B200 decoder is batched: decode_batch() is the native call and decode() is a batch of one.
// Like the reference's handle (c_api/decoder.rs:120, `&mut`), one decoder object must not be used
// from two threads at once.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "host.hpp"

namespace ldpc {

struct DecoderOutput {
    std::vector<uint8_t> codeword;   // n hard bits, one 0/1 per byte
    size_t iterations = 0;
    bool success = false;            // Ok(..) vs Err(..) of the reference
};

struct BatchStats {                  // filled by every decode_batch call (device-side timing hooks)
    float ingest_ms = 0, decode_ms = 0, emit_ms = 0;
    long long kernel_launches = 0;
};

class LdpcDecoder {
public:
    virtual ~LdpcDecoder() = default;
    // reference semantics; llrs.size() must equal n (punctured length when a puncturer is attached)
    virtual bool decode(const double* llrs, size_t llrs_len, size_t max_iterations, DecoderOutput* out) = 0;
    // Batched decode on host buffers.  llrs: [nframes][llrs_len]; out: [nframes][out_stride] bytes of
    // which the first out_len are written; iterations[f] >= 0 on success, -1 on failure.
    virtual bool decode_batch(const void* llrs, bool is_f64, size_t llrs_len, size_t nframes, uint32_t max_iterations,
                              uint8_t* out, size_t out_len, size_t out_stride, int32_t* iterations) = 0;
    // Asynchronous form of decode_batch: returns a ticket (>= 0) as soon as the work is enqueued, -1 on an
    // argument / CUDA error.  The caller's buffers must stay valid (and, to overlap, pinned) until wait(ticket)
    // returns; tickets complete in submission order.  decode_batch = submit_batch + wait.
    virtual int64_t submit_batch(const void* llrs, bool is_f64, size_t llrs_len, size_t nframes, uint32_t max_iterations,
                                 uint8_t* out, size_t out_len, size_t out_stride, int32_t* iterations) = 0;
    virtual bool wait(int64_t ticket) = 0;
    // Test hook: decode_batch that also returns the posterior LLRs [nframes][n] as f64 (float decoders only).
    virtual bool decode_batch_posteriors(const void* llrs, bool is_f64, size_t llrs_len, size_t nframes, uint32_t max_iterations,
                                         uint8_t* out, size_t out_len, size_t out_stride, int32_t* iterations, double* posteriors) = 0;
    // Same on device buffers (inputs already resident in HBM), asynchronous on `stream`.
    virtual bool decode_batch_device(const void* d_llrs, bool is_f64, size_t llrs_len, size_t nframes,
                                     uint32_t max_iterations, uint8_t* d_out, size_t out_len, size_t out_stride,
                                     int32_t* d_iterations, cudaStream_t stream) = 0;
    // The same with an explicit workspace lane (0 or 1): two calls on different lanes and streams may be in flight
    // at once (the BER engine keeps two batches resident); decode_batch_device is lane 0.
    virtual bool decode_batch_device_lane(int lane, const void* d_llrs, bool is_f64, size_t llrs_len, size_t nframes,
                                          uint32_t max_iterations, uint8_t* d_out, size_t out_len, size_t out_stride,
                                          int32_t* d_iterations, cudaStream_t stream) = 0;
    virtual int n() const = 0;
    virtual int k() const = 0;
    virtual int edges() const = 0;
    virtual size_t expected_llrs_len() const = 0;
    virtual const BatchStats& stats() const = 0;
};

struct DecoderOptions {
    int device = -1;                 // -1: current device
    int max_tiles = 0;               // frames per launch / 128; 0: automatic (whole waves of CTAs that fit in free HBM)
    int words_per_lane = 0;          // int8 decoders: 1 = 128-frame tiles, 4 = 512-frame tiles, 0 = by batch size
    int layered_path = 0;            // layered decoders: 0 = automatic, 1 = frame-interleaved tiles (K3), 2 = frame per CTA (K3q)
};

// DecoderFactory::build_decoder.  Returns nullptr and sets last_error() on failure.
std::unique_ptr<LdpcDecoder> build_decoder(const DecoderImplementation& impl, const Graph& h,
                                           const Puncturer* puncturer, const DecoderOptions& opt);

}  // namespace ldpc
