// ldpc_toolbox_b200/csrc/ber.hpp — on-device BER Monte-Carlo engine (one Eb/N0 point at a time).
// Mirrors BerTest / Worker of the reference (src/simulation/ber.rs:246-282, :297-368, :436-481):
// the host keeps the stop rule and the statistics, the device does everything per frame.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <memory>
#include <string>

#include "decoder.hpp"
#include "host.hpp"

namespace ldpc {

// counters: frames, bit_errors, frame_errors, false_decodes, total_iterations, correct_iterations,
//           bch_bit_errors, bch_frame_errors, bch_correct_iterations          (ber.rs:313-337)
constexpr int kBerCounters = 9;

class BerEngine {
public:
    static std::unique_ptr<BerEngine> create(const Graph& g, const DecoderImplementation& impl, const Puncturer* punct,
                                             const DecoderOptions& opt);
    ~BerEngine();
    // Simulates global frames [first_frame, first_frame + nframes) at ebn0_db and ADDS to counters[kBerCounters].
    // Optional host dump buffers (test hooks): LLRs [nframes][n_tx] f32, decoded info bytes [nframes][k],
    // iterations [nframes], messages [nframes][ceil(k/32)].
    bool run(float ebn0_db, uint32_t max_iterations, uint64_t first_frame, uint64_t nframes, uint64_t seed,
             uint64_t bch_max_errors, uint64_t* counters, float* dump_llrs = nullptr, uint8_t* dump_decoded = nullptr,
             int32_t* dump_iters = nullptr, uint32_t* dump_messages = nullptr);
    // Asynchronous form: submit() enqueues front-end, decode, back-end and the read-back of the nine counters on one
    // of two lanes (own stream, buffers and decoder workspace) and returns a ticket (>= 0, -1 on error) without
    // waiting; wait(ticket) blocks until that batch is done and ADDS its counters.  Two tickets may be in flight,
    // so the host's stop rule / counter reduction for batch i overlaps the kernels of batch i+1 (ber.rs:312-343:
    // the reference's controller also keeps receiving results while the workers run ahead).
    int64_t submit(float ebn0_db, uint32_t max_iterations, uint64_t first_frame, uint64_t nframes, uint64_t seed,
                   uint64_t bch_max_errors);
    bool wait(int64_t ticket, uint64_t* counters);
    double noise_sigma(float ebn0_db) const;
    // "BPSK" or "8PSK" (reference src/simulation/factory.rs:56-86); interleaving_columns: 0 = none, n = DVB-S2
    // bit interleaver with n columns, -n = rows read backwards (reference src/simulation/ber.rs:250-252)
    bool set_modulation(const std::string& name, int interleaving_columns);
    int n() const { return n_; }
    int k() const { return k_; }
    int n_tx() const { return n_tx_; }       // transmitted symbols per frame ("Frame size (N)")
    double rate() const { return rate_; }
    long long kernel_launches() const { return launches_; }
    LdpcDecoder* decoder() { return decoder_.get(); }

private:
    BerEngine() = default;
    struct Lane {
        float* d_llrs = nullptr; uint32_t* d_messages = nullptr; uint8_t* d_decoded = nullptr; int32_t* d_iters = nullptr;
        unsigned long long* d_counters = nullptr;
        unsigned long long* h_counters = nullptr;      // pinned
        size_t cap_frames = 0;
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
        int64_t ticket = -1;                           // ticket in flight on this lane, or -1
    };
    bool ensure(Lane& ln, size_t nframes);
    bool enqueue(Lane& ln, int lane_index, float ebn0_db, uint32_t max_iterations, uint64_t first_frame, uint64_t nframes, uint64_t seed,
                 uint64_t bch_max_errors);
    Lane lanes_[2];
    int64_t next_ticket_ = 0;
    EncoderPlan plan_;
    std::unique_ptr<LdpcDecoder> decoder_;
    int n_ = 0, m_ = 0, k_ = 0, n_tx_ = 0, device_ = 0, g0_words_ = 0;
    int modulation_ = 0, il_cols_ = 0, il_backwards_ = 0;
    double rate_ = 0;
    int* d_h0_ptr_ = nullptr; int* d_h0_idx_ = nullptr; uint32_t* d_g0_ = nullptr; int* d_kept_ = nullptr;
    long long launches_ = 0;
};

}  // namespace ldpc
