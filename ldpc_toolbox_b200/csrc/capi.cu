// ldpc_toolbox_b200/csrc/capi.cu — extern "C" boundary declared in include/ldpc_toolbox.h.
// Replaces reference src/c_api.rs, src/c_api/decoder.rs and src/c_api/encoder.rs.
#include <cstdio>
#include <cstring>
#include <exception>
#include <fstream>
#include <sstream>

#include "../../include/ldpc_toolbox.h"
#include "ber.hpp"
#include "decoder.hpp"
#include "decoder_impl.hpp"

namespace ldpc {
void resolve_decoder_stats(LdpcDecoder* d);
float average_decode_ms(LdpcDecoder* d, int64_t* launches);
}

using namespace ldpc;

namespace {

// No C++ exception may cross the C boundary (std::bad_alloc on an absurd alist header, std::length_error ...):
// the call fails with its error value and ldpc_toolbox_last_error() says why.
template <class R, class F>
R guarded(R on_error, F&& f) noexcept {
    try { return f(); }
    catch (const std::exception& e) { set_last_error(std::string("internal error: ") + e.what()); }
    catch (...) { set_last_error("internal error"); }
    return on_error;
}

struct DecoderHandle {                       // reference c_api/decoder.rs:19-23
    std::unique_ptr<LdpcDecoder> decoder;
};

struct EncoderHandle {                       // reference c_api/encoder.rs:14-18
    EncoderPlan plan;
    bool punctured = false;
    std::vector<int> kept;                   // transmitted positions when punctured
    std::vector<uint8_t> msg, cw;
};

bool slurp(const char* path, std::string* out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { set_last_error(std::string("cannot read ") + path); return false; }
    std::ostringstream ss;
    ss << f.rdbuf();
    *out = ss.str();
    return true;
}

bool make_puncturer(const char* puncturing, std::unique_ptr<Puncturer>* p) {
    if (puncturing[0] == '\0') return true;                 // c_api/decoder.rs:28-33
    std::vector<bool> pat;
    if (!parse_puncturing_pattern(puncturing, &pat)) { set_last_error("invalid puncturing pattern"); return false; }
    *p = std::make_unique<Puncturer>(pat);
    return true;
}

void* decoder_new(const std::string& alist, const char* implementation, const char* puncturing, int device, int max_tiles) {
    return guarded<void*>(nullptr, [&]() -> void* {
        Graph g;
        std::string err;
        if (!Graph::from_alist(alist, &g, &err)) { set_last_error(err); return nullptr; }
        DecoderImplementation impl;
        if (!DecoderImplementation::parse(implementation, &impl)) { set_last_error("invalid decoder implementation"); return nullptr; }
        std::unique_ptr<Puncturer> p;
        if (!make_puncturer(puncturing, &p)) return nullptr;
        DecoderOptions opt;
        opt.device = device;
        opt.max_tiles = max_tiles;
        auto dec = build_decoder(impl, g, p.get(), opt);
        if (!dec) return nullptr;
        auto* h = new DecoderHandle();
        h->decoder = std::move(dec);
        return h;
    });
}

void* encoder_new(const std::string& alist, const char* puncturing) {
    return guarded<void*>(nullptr, [&]() -> void* {
        Graph g;
        std::string err;
        if (!Graph::from_alist(alist, &g, &err)) { set_last_error(err); return nullptr; }
        std::unique_ptr<Puncturer> p;
        if (!make_puncturer(puncturing, &p)) return nullptr;
        auto h = std::make_unique<EncoderHandle>();
        if (!EncoderPlan::from_graph(g, &h->plan, &err)) { set_last_error(err); return nullptr; }
        if (p) {
            h->punctured = true;
            // the reference only fails at encode time (puncture() -> unwrap); remember an empty map
            if (!p->puncture_map((size_t)g.n, &h->kept)) h->kept.assign(1, -1);
        }
        h->msg.resize((size_t)h->plan.k);
        h->cw.resize((size_t)h->plan.n);
        return h.release();
    });
}

template <class T>
int32_t decode_single(void* decoder, uint8_t* output, size_t output_len, const T* llrs, size_t llrs_len, uint32_t max_it) {
    return guarded<int32_t>(-2, [&]() -> int32_t {
        if (!decoder || (!output && output_len) || !llrs) return -2;
        auto* h = static_cast<DecoderHandle*>(decoder);
        int32_t it = -2;
        if (!h->decoder->decode_batch(llrs, sizeof(T) == 8, llrs_len, 1, max_it, output, output_len, output_len, &it)) return -2;
        return it;
    });
}

template <class T>
int32_t decode_batch(void* decoder, uint8_t* output, size_t output_len, size_t output_stride, const T* llrs, size_t llrs_len,
                     size_t nframes, uint32_t max_it, int32_t* iterations) {
    return guarded<int32_t>(-2, [&]() -> int32_t {
        if (!decoder || !iterations || !llrs || (!output && output_len)) return -2;
        auto* h = static_cast<DecoderHandle*>(decoder);
        return h->decoder->decode_batch(llrs, sizeof(T) == 8, llrs_len, nframes, max_it, output, output_len, output_stride, iterations) ? 0 : -2;
    });
}

template <class T>
int64_t submit_batch(void* decoder, uint8_t* output, size_t output_len, size_t output_stride, const T* llrs, size_t llrs_len,
                     size_t nframes, uint32_t max_it, int32_t* iterations) {
    return guarded<int64_t>(-2, [&]() -> int64_t {
        if (!decoder || !iterations || !llrs || (!output && output_len)) return -2;
        auto* h = static_cast<DecoderHandle*>(decoder);
        const int64_t t = h->decoder->submit_batch(llrs, sizeof(T) == 8, llrs_len, nframes, max_it, output, output_len, output_stride, iterations);
        return t >= 0 ? t : -2;
    });
}

template <class T>
int32_t decode_batch_device(void* decoder, uint8_t* d_output, size_t output_len, size_t output_stride, const T* d_llrs,
                            size_t llrs_len, size_t nframes, uint32_t max_it, int32_t* d_iterations, void* stream) {
    return guarded<int32_t>(-2, [&]() -> int32_t {
        if (!decoder || !d_iterations || !d_llrs || (!d_output && output_len)) return -2;
        auto* h = static_cast<DecoderHandle*>(decoder);
        return h->decoder->decode_batch_device(d_llrs, sizeof(T) == 8, llrs_len, nframes, max_it, d_output, output_len, output_stride,
                                               d_iterations, static_cast<cudaStream_t>(stream)) ? 0 : -2;
    });
}

}  // namespace

extern "C" {

void* ldpc_toolbox_decoder_ctor(const char* alist_file_path, const char* implementation, const char* puncturing) {
    return guarded<void*>(nullptr, [&]() -> void* {
        if (!alist_file_path || !implementation || !puncturing) return nullptr;
        std::string s;
        if (!slurp(alist_file_path, &s)) return nullptr;
        return decoder_new(s, implementation, puncturing, -1, 0);
    });
}

void* ldpc_toolbox_decoder_ctor_alist_string(const char* alist, const char* implementation, const char* puncturing) {
    return guarded<void*>(nullptr, [&]() -> void* {
        if (!alist || !implementation || !puncturing) return nullptr;
        return decoder_new(alist, implementation, puncturing, -1, 0);
    });
}

void* ldpc_toolbox_decoder_ctor_ex(const char* alist, int alist_is_path, const char* implementation, const char* puncturing,
                                   int device, int max_tiles) {
    return guarded<void*>(nullptr, [&]() -> void* {
        if (!alist || !implementation || !puncturing) return nullptr;
        if (alist_is_path) {
            std::string s;
            if (!slurp(alist, &s)) return nullptr;
            return decoder_new(s, implementation, puncturing, device, max_tiles);
        }
        return decoder_new(alist, implementation, puncturing, device, max_tiles);
    });
}

void ldpc_toolbox_decoder_dtor(void* decoder) { delete static_cast<DecoderHandle*>(decoder); }

int32_t ldpc_toolbox_decoder_decode_f64(void* decoder, uint8_t* output, size_t output_len, const double* llrs, size_t llrs_len,
                                        uint32_t max_iterations) {
    return decode_single(decoder, output, output_len, llrs, llrs_len, max_iterations);
}

int32_t ldpc_toolbox_decoder_decode_f32(void* decoder, uint8_t* output, size_t output_len, const float* llrs, size_t llrs_len,
                                        uint32_t max_iterations) {
    return decode_single(decoder, output, output_len, llrs, llrs_len, max_iterations);
}

int32_t ldpc_toolbox_decoder_decode_batch_f32(void* decoder, uint8_t* output, size_t output_len, size_t output_stride,
                                              const float* llrs, size_t llrs_len, size_t nframes, uint32_t max_iterations,
                                              int32_t* iterations) {
    return decode_batch(decoder, output, output_len, output_stride, llrs, llrs_len, nframes, max_iterations, iterations);
}

int32_t ldpc_toolbox_decoder_decode_batch_f64(void* decoder, uint8_t* output, size_t output_len, size_t output_stride,
                                              const double* llrs, size_t llrs_len, size_t nframes, uint32_t max_iterations,
                                              int32_t* iterations) {
    return decode_batch(decoder, output, output_len, output_stride, llrs, llrs_len, nframes, max_iterations, iterations);
}

int64_t ldpc_toolbox_decoder_submit_batch_f32(void* decoder, uint8_t* output, size_t output_len, size_t output_stride,
                                              const float* llrs, size_t llrs_len, size_t nframes, uint32_t max_iterations,
                                              int32_t* iterations) {
    return submit_batch(decoder, output, output_len, output_stride, llrs, llrs_len, nframes, max_iterations, iterations);
}

int64_t ldpc_toolbox_decoder_submit_batch_f64(void* decoder, uint8_t* output, size_t output_len, size_t output_stride,
                                              const double* llrs, size_t llrs_len, size_t nframes, uint32_t max_iterations,
                                              int32_t* iterations) {
    return submit_batch(decoder, output, output_len, output_stride, llrs, llrs_len, nframes, max_iterations, iterations);
}

int32_t ldpc_toolbox_decoder_decode_batch_posteriors_f32(void* decoder, uint8_t* output, size_t output_len, size_t output_stride,
                                                         const float* llrs, size_t llrs_len, size_t nframes, uint32_t max_iterations,
                                                         int32_t* iterations, double* posteriors) {
    return guarded<int32_t>(-2, [&]() -> int32_t {
        if (!decoder || !iterations || !llrs || !posteriors || (!output && output_len)) return -2;
        return static_cast<DecoderHandle*>(decoder)->decoder->decode_batch_posteriors(llrs, false, llrs_len, nframes, max_iterations, output,
                                                                                      output_len, output_stride, iterations, posteriors) ? 0 : -2;
    });
}

int32_t ldpc_toolbox_decoder_decode_batch_posteriors_f64(void* decoder, uint8_t* output, size_t output_len, size_t output_stride,
                                                         const double* llrs, size_t llrs_len, size_t nframes, uint32_t max_iterations,
                                                         int32_t* iterations, double* posteriors) {
    return guarded<int32_t>(-2, [&]() -> int32_t {
        if (!decoder || !iterations || !llrs || !posteriors || (!output && output_len)) return -2;
        return static_cast<DecoderHandle*>(decoder)->decoder->decode_batch_posteriors(llrs, true, llrs_len, nframes, max_iterations, output,
                                                                                      output_len, output_stride, iterations, posteriors) ? 0 : -2;
    });
}

int32_t ldpc_toolbox_decoder_wait(void* decoder, int64_t ticket) {
    return guarded<int32_t>(-2, [&]() -> int32_t {
        if (!decoder || ticket < 0) return -2;
        return static_cast<DecoderHandle*>(decoder)->decoder->wait(ticket) ? 0 : -2;
    });
}

int32_t ldpc_toolbox_decoder_decode_batch_device_f32(void* decoder, uint8_t* d_output, size_t output_len, size_t output_stride,
                                                     const float* d_llrs, size_t llrs_len, size_t nframes,
                                                     uint32_t max_iterations, int32_t* d_iterations, void* cuda_stream) {
    return decode_batch_device(decoder, d_output, output_len, output_stride, d_llrs, llrs_len, nframes, max_iterations,
                               d_iterations, cuda_stream);
}

int32_t ldpc_toolbox_decoder_decode_batch_device_f64(void* decoder, uint8_t* d_output, size_t output_len, size_t output_stride,
                                                     const double* d_llrs, size_t llrs_len, size_t nframes,
                                                     uint32_t max_iterations, int32_t* d_iterations, void* cuda_stream) {
    return decode_batch_device(decoder, d_output, output_len, output_stride, d_llrs, llrs_len, nframes, max_iterations,
                               d_iterations, cuda_stream);
}

size_t ldpc_toolbox_decoder_codeword_len(void* d) { return d ? (size_t) static_cast<DecoderHandle*>(d)->decoder->n() : 0; }
size_t ldpc_toolbox_decoder_info_len(void* d) { return d ? (size_t) static_cast<DecoderHandle*>(d)->decoder->k() : 0; }
size_t ldpc_toolbox_decoder_num_edges(void* d) { return d ? (size_t) static_cast<DecoderHandle*>(d)->decoder->edges() : 0; }
size_t ldpc_toolbox_decoder_llrs_len(void* d) { return d ? static_cast<DecoderHandle*>(d)->decoder->expected_llrs_len() : 0; }

int64_t ldpc_toolbox_decoder_last_timing(void* d, float* ms3) {
    if (!d) return 0;
    auto* dec = static_cast<DecoderHandle*>(d)->decoder.get();
    resolve_decoder_stats(dec);
    const BatchStats& s = dec->stats();
    if (ms3) { ms3[0] = s.ingest_ms; ms3[1] = s.decode_ms; ms3[2] = s.emit_ms; }
    return s.kernel_launches;
}

float ldpc_toolbox_decoder_average_decode_ms(void* d, int64_t* launches) {
    if (launches) *launches = 0;
    if (!d) return 0.0f;
    return average_decode_ms(static_cast<DecoderHandle*>(d)->decoder.get(), launches);
}

void* ldpc_toolbox_encoder_ctor(const char* alist_file_path, const char* puncturing) {
    return guarded<void*>(nullptr, [&]() -> void* {
        if (!alist_file_path || !puncturing) return nullptr;
        std::string s;
        if (!slurp(alist_file_path, &s)) return nullptr;
        return encoder_new(s, puncturing);
    });
}

void* ldpc_toolbox_encoder_ctor_alist_string(const char* alist, const char* puncturing) {
    return guarded<void*>(nullptr, [&]() -> void* {
        if (!alist || !puncturing) return nullptr;
        return encoder_new(alist, puncturing);
    });
}

void ldpc_toolbox_encoder_dtor(void* encoder) { delete static_cast<EncoderHandle*>(encoder); }

void ldpc_toolbox_encoder_encode(void* encoder, uint8_t* output, size_t output_len, const uint8_t* input, size_t input_len) {
    if (!encoder || !output || !input) return;
    auto* h = static_cast<EncoderHandle*>(encoder);
    // the reference panics on a length mismatch (ndarray dot / assert_eq!, c_api/encoder.rs:44-48)
    if (input_len != (size_t)h->plan.k) { set_last_error("encoder: input_len != k"); return; }
    const size_t tx_len = h->punctured ? h->kept.size() : (size_t)h->plan.n;
    if ((h->punctured && !h->kept.empty() && h->kept[0] < 0) || output_len != tx_len) {
        set_last_error("encoder: output_len does not match the (punctured) codeword length");
        return;
    }
    for (size_t i = 0; i < input_len; ++i) h->msg[i] = input[i] == 1 ? 1 : 0;    // c_api/encoder.rs:38-42
    h->plan.encode(h->msg.data(), h->cw.data());
    if (h->punctured) for (size_t i = 0; i < tx_len; ++i) output[i] = h->cw[(size_t)h->kept[i]];
    else memcpy(output, h->cw.data(), tx_len);
}

void* ldpc_toolbox_ber_ctor(const char* alist, int alist_is_path, const char* implementation, const char* puncturing, int device,
                            int max_tiles) {
    return guarded<void*>(nullptr, [&]() -> void* {
        if (!alist || !implementation || !puncturing) return nullptr;
        std::string text;
        if (alist_is_path) { if (!slurp(alist, &text)) return nullptr; } else text = alist;
        Graph g;
        std::string err;
        if (!Graph::from_alist(text, &g, &err)) { set_last_error(err); return nullptr; }
        DecoderImplementation impl;
        if (!DecoderImplementation::parse(implementation, &impl)) { set_last_error("invalid decoder implementation"); return nullptr; }
        std::unique_ptr<Puncturer> p;
        if (!make_puncturer(puncturing, &p)) return nullptr;
        DecoderOptions opt;
        opt.device = device;
        opt.max_tiles = max_tiles;
        if (device >= 0 && cudaSetDevice(device) != cudaSuccess) { set_last_error("cudaSetDevice failed"); return nullptr; }
        return BerEngine::create(g, impl, p.get(), opt).release();
    });
}

void ldpc_toolbox_ber_dtor(void* ber) { delete static_cast<BerEngine*>(ber); }

int32_t ldpc_toolbox_ber_set_modulation(void* ber, const char* modulation, int32_t interleaving_columns) {
    return guarded<int32_t>(-2, [&]() -> int32_t {
        if (!ber || !modulation) return -2;
        return static_cast<BerEngine*>(ber)->set_modulation(modulation, interleaving_columns) ? 0 : -2;
    });
}

int32_t ldpc_toolbox_ber_run(void* ber, float ebn0_db, uint32_t max_iterations, uint64_t first_frame, uint64_t nframes, uint64_t seed,
                             uint64_t bch_max_errors, uint64_t* counters) {
    return guarded<int32_t>(-2, [&]() -> int32_t {
        if (!ber || !counters) return -2;
        return static_cast<BerEngine*>(ber)->run(ebn0_db, max_iterations, first_frame, nframes, seed, bch_max_errors, counters) ? 0 : -2;
    });
}

int32_t ldpc_toolbox_ber_run_dump(void* ber, float ebn0_db, uint32_t max_iterations, uint64_t first_frame, uint64_t nframes,
                                  uint64_t seed, uint64_t bch_max_errors, uint64_t* counters, float* llrs, uint8_t* decoded,
                                  int32_t* iterations, uint32_t* messages) {
    return guarded<int32_t>(-2, [&]() -> int32_t {
        if (!ber || !counters) return -2;
        return static_cast<BerEngine*>(ber)->run(ebn0_db, max_iterations, first_frame, nframes, seed, bch_max_errors, counters, llrs,
                                                 decoded, iterations, messages) ? 0 : -2;
    });
}

int64_t ldpc_toolbox_ber_submit(void* ber, float ebn0_db, uint32_t max_iterations, uint64_t first_frame, uint64_t nframes, uint64_t seed,
                                uint64_t bch_max_errors) {
    return guarded<int64_t>(-2, [&]() -> int64_t {
        if (!ber) return -2;
        const int64_t t = static_cast<BerEngine*>(ber)->submit(ebn0_db, max_iterations, first_frame, nframes, seed, bch_max_errors);
        return t >= 0 ? t : -2;
    });
}

int32_t ldpc_toolbox_ber_wait(void* ber, int64_t ticket, uint64_t* counters) {
    return guarded<int32_t>(-2, [&]() -> int32_t {
        if (!ber || !counters) return -2;
        return static_cast<BerEngine*>(ber)->wait(ticket, counters) ? 0 : -2;
    });
}

void ldpc_toolbox_ber_dims(void* ber, uint64_t* what3) {
    if (!ber || !what3) return;
    auto* b = static_cast<BerEngine*>(ber);
    what3[0] = (uint64_t)b->k(); what3[1] = (uint64_t)b->n(); what3[2] = (uint64_t)b->n_tx();
}
double ldpc_toolbox_ber_rate(void* ber) { return ber ? static_cast<BerEngine*>(ber)->rate() : 0.0; }
double ldpc_toolbox_ber_noise_sigma(void* ber, float ebn0_db) { return ber ? static_cast<BerEngine*>(ber)->noise_sigma(ebn0_db) : 0.0; }

const char* ldpc_toolbox_last_error(void) { return last_error().c_str(); }

int32_t ldpc_toolbox_num_implementations(void) { return (int32_t)DecoderImplementation::all_names().size(); }
const char* ldpc_toolbox_implementation_name(int32_t index) {
    const auto& n = DecoderImplementation::all_names();
    return index >= 0 && (size_t)index < n.size() ? n[(size_t)index].c_str() : nullptr;
}

}  // extern "C"
