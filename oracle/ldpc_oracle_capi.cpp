// oracle/ldpc_oracle_capi.cpp — TEST INFRASTRUCTURE ONLY.
//
// C ABI over the CPU restatement, shaped like the reference's FFI
// (src/c_api/decoder.rs, src/c_api/encoder.rs, include/ldpc_toolbox.h) but with
// the prefix `ldpc_oracle_` so the product library and the checker can live in
// one process without symbol clashes.  Extra entry points (batch decode with
// worker threads, posterior read-back, BER run) exist only for tests and for the
// cpu_baseline leg of bench.py.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <thread>

#include "ldpc_oracle.hpp"

using namespace oracle;

namespace {
struct DecoderHandle {                      // src/c_api/decoder.rs:19-23
    SparseMatrix h;
    std::string implementation;
    std::unique_ptr<LdpcDecoder> decoder;
    std::unique_ptr<Puncturer> puncturer;
    bool linear = false;
};
struct EncoderHandle {                      // src/c_api/encoder.rs:14-18
    std::unique_ptr<Encoder> encoder;
    std::unique_ptr<Puncturer> puncturer;
};

bool read_file(const char* path, std::string* out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    *out = ss.str();
    return true;
}

DecoderHandle* make_decoder(const std::string& alist, const char* impl, const char* punct) {   // decoder.rs:26-36
    auto d = std::make_unique<DecoderHandle>();
    std::string err;
    if (!SparseMatrix::from_alist(alist, &d->h, &err)) return nullptr;
    d->implementation = impl;
    d->decoder = build_decoder(d->implementation, d->h);
    if (!d->decoder) return nullptr;
    if (punct[0] != '\0') {
        std::vector<bool> pat;
        if (!parse_puncturing_pattern(punct, &pat)) return nullptr;
        d->puncturer = std::make_unique<Puncturer>(pat);
    }
    return d.release();
}

EncoderHandle* make_encoder(const std::string& alist, const char* punct) {                      // encoder.rs:21-32
    auto e = std::make_unique<EncoderHandle>();
    SparseMatrix h;
    std::string err;
    if (!SparseMatrix::from_alist(alist, &h, &err)) return nullptr;
    if (punct[0] != '\0') {
        std::vector<bool> pat;
        if (!parse_puncturing_pattern(punct, &pat)) return nullptr;
        e->puncturer = std::make_unique<Puncturer>(pat);
    }
    e->encoder = Encoder::from_h(h, &err);
    if (!e->encoder) return nullptr;
    return e.release();
}

// decoder.rs:50-67.  Returns iterations, -1 on decode failure, -2 where the reference panics.
int32_t decode_one(LdpcDecoder* dec, const Puncturer* p, uint8_t* out, size_t out_len, const double* llrs,
                   size_t llrs_len, uint32_t max_it) {
    std::vector<double> dep;
    if (p) {
        if (!p->depuncture(llrs, llrs_len, &dep)) return -2;
        llrs = dep.data();
        llrs_len = dep.size();
    }
    DecodeResult r = dec->decode(llrs, llrs_len, max_it);
    if (r.error || out_len > r.codeword.size()) return -2;
    std::memcpy(out, r.codeword.data(), out_len);
    return r.success ? (int32_t)r.iterations : -1;
}
}  // namespace

extern "C" {

void* ldpc_oracle_decoder_ctor(const char* path, const char* impl, const char* punct) {
    std::string s;
    if (!read_file(path, &s)) return nullptr;
    return make_decoder(s, impl, punct);
}
void* ldpc_oracle_decoder_ctor_alist_string(const char* alist, const char* impl, const char* punct) {
    return make_decoder(alist, impl, punct);
}
void ldpc_oracle_decoder_dtor(void* d) { delete static_cast<DecoderHandle*>(d); }

int32_t ldpc_oracle_decoder_decode_f64(void* d, uint8_t* out, size_t out_len, const double* llrs, size_t llrs_len,
                                       uint32_t max_it) {
    auto* h = static_cast<DecoderHandle*>(d);
    return decode_one(h->decoder.get(), h->puncturer.get(), out, out_len, llrs, llrs_len, max_it);
}
int32_t ldpc_oracle_decoder_decode_f32(void* d, uint8_t* out, size_t out_len, const float* llrs, size_t llrs_len,
                                       uint32_t max_it) {                                       // decoder.rs:69-72
    std::vector<double> w(llrs, llrs + llrs_len);
    return ldpc_oracle_decoder_decode_f64(d, out, out_len, w.data(), llrs_len, max_it);
}

void ldpc_oracle_decoder_set_linear_search(void* d, int on) {
    auto* h = static_cast<DecoderHandle*>(d);
    h->linear = on != 0;
    h->decoder->set_linear_search_send(h->linear);
}

size_t ldpc_oracle_decoder_n(void* d) { return static_cast<DecoderHandle*>(d)->h.num_cols(); }
size_t ldpc_oracle_decoder_m(void* d) { return static_cast<DecoderHandle*>(d)->h.num_rows(); }
size_t ldpc_oracle_decoder_edges(void* d) { return static_cast<DecoderHandle*>(d)->h.nnz(); }

// posterior LLRs (as f64) of the last single-frame decode; returns count written
size_t ldpc_oracle_decoder_posteriors(void* d, double* out, size_t cap) {
    auto p = static_cast<DecoderHandle*>(d)->decoder->posteriors();
    size_t c = std::min(cap, p.size());
    std::memcpy(out, p.data(), c * sizeof(double));
    return c;
}

// Frame-parallel batch decode (one private decoder per worker thread, like the BER
// workers of src/simulation/ber.rs:370-392).  iterations[f] = count, -1 failure, -2 panic.
static int32_t batch_impl(DecoderHandle* h, uint8_t* out, size_t out_len, const void* llrs, bool is_f32,
                          size_t llrs_len, size_t nframes, uint32_t max_it, int32_t* iterations, int nthreads) {
    if (nthreads < 1) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
    nthreads = (int)std::min<size_t>((size_t)nthreads, std::max<size_t>(nframes, 1));
    auto work = [&](int tid) {
        std::unique_ptr<LdpcDecoder> own;
        LdpcDecoder* dec = h->decoder.get();
        if (tid != 0) { own = build_decoder(h->implementation, h->h); own->set_linear_search_send(h->linear); dec = own.get(); }
        std::vector<double> w(llrs_len);
        for (size_t f = (size_t)tid; f < nframes; f += (size_t)nthreads) {
            if (is_f32) { const float* p = (const float*)llrs + f * llrs_len; for (size_t i = 0; i < llrs_len; ++i) w[i] = (double)p[i]; }
            else std::memcpy(w.data(), (const double*)llrs + f * llrs_len, llrs_len * sizeof(double));
            iterations[f] = decode_one(dec, h->puncturer.get(), out + f * out_len, out_len, w.data(), llrs_len, max_it);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    return 0;
}
int32_t ldpc_oracle_decoder_decode_batch_f32(void* d, uint8_t* out, size_t out_len, const float* llrs, size_t llrs_len,
                                             size_t nframes, uint32_t max_it, int32_t* iterations, int nthreads) {
    return batch_impl(static_cast<DecoderHandle*>(d), out, out_len, llrs, true, llrs_len, nframes, max_it, iterations, nthreads);
}
int32_t ldpc_oracle_decoder_decode_batch_f64(void* d, uint8_t* out, size_t out_len, const double* llrs, size_t llrs_len,
                                             size_t nframes, uint32_t max_it, int32_t* iterations, int nthreads) {
    return batch_impl(static_cast<DecoderHandle*>(d), out, out_len, llrs, false, llrs_len, nframes, max_it, iterations, nthreads);
}

void* ldpc_oracle_encoder_ctor(const char* path, const char* punct) {
    std::string s;
    if (!read_file(path, &s)) return nullptr;
    return make_encoder(s, punct);
}
void* ldpc_oracle_encoder_ctor_alist_string(const char* alist, const char* punct) { return make_encoder(alist, punct); }
void ldpc_oracle_encoder_dtor(void* e) { delete static_cast<EncoderHandle*>(e); }
int ldpc_oracle_encoder_is_staircase(void* e) { return static_cast<EncoderHandle*>(e)->encoder->is_staircase() ? 1 : 0; }

// encoder.rs:37-52.  Returns 0, or -2 where the reference panics (length mismatch).
int32_t ldpc_oracle_encoder_encode(void* e, uint8_t* out, size_t out_len, const uint8_t* in, size_t in_len) {
    auto* h = static_cast<EncoderHandle*>(e);
    const Encoder& enc = *h->encoder;
    if (in_len != enc.k()) return -2;
    std::vector<uint8_t> msg(in_len), cw(enc.n());
    for (size_t i = 0; i < in_len; ++i) msg[i] = in[i] == 1 ? 1 : 0;
    enc.encode(msg.data(), cw.data());
    std::vector<uint8_t> tx;
    if (h->puncturer) { if (!h->puncturer->puncture(cw, &tx)) return -2; } else tx = cw;
    if (tx.size() != out_len) return -2;
    std::memcpy(out, tx.data(), out_len);
    return 0;
}

int ldpc_oracle_num_implementations() { return (int)implementation_names().size(); }
const char* ldpc_oracle_implementation_name(int i) { return implementation_names()[(size_t)i].c_str(); }

// alist round trip (src/sparse.rs:250-341): parse then write; returns bytes needed (incl. NUL) or 0 on parse error
size_t ldpc_oracle_alist_roundtrip(const char* alist, int padding, char* out, size_t cap) {
    SparseMatrix h;
    std::string err;
    if (!SparseMatrix::from_alist(alist, &h, &err)) return 0;
    std::string s = h.alist(padding != 0);
    if (out && cap > s.size()) std::memcpy(out, s.c_str(), s.size() + 1);
    return s.size() + 1;
}

// counters[0..5] = frames, bit_errors, frame_errors, false_decodes, total_iterations, correct_iterations
int32_t ldpc_oracle_ber_run(const char* alist, const char* impl, const char* punct, float ebn0_db, uint32_t max_it,
                            uint64_t frames, uint64_t max_frame_errors, int nthreads, uint64_t seed, int linear_search,
                            uint64_t* counters, double* elapsed_s) {
    SparseMatrix h;
    std::string err;
    if (!SparseMatrix::from_alist(alist, &h, &err)) return -1;
    BerCounters c;
    if (!ber_run(h, impl, punct, ebn0_db, max_it, frames, max_frame_errors, nthreads, seed, linear_search != 0, &c, &err)) {
        std::fprintf(stderr, "ldpc_oracle_ber_run: %s\n", err.c_str());
        return -1;
    }
    counters[0] = c.num_frames; counters[1] = c.bit_errors; counters[2] = c.frame_errors;
    counters[3] = c.false_decodes; counters[4] = c.total_iterations; counters[5] = c.correct_iterations;
    *elapsed_s = c.elapsed_s;
    return 0;
}

double ldpc_oracle_noise_sigma(double rate, double bits_per_symbol, float ebn0_db) {
    return noise_sigma(rate, bits_per_symbol, ebn0_db);
}

}  // extern "C"
