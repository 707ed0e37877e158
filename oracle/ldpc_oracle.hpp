// oracle/ldpc_oracle.hpp
//
// TEST INFRASTRUCTURE ONLY.  CPU restatement (C++17) of the hot path of
// daniestevez/ldpc-toolbox v0.12.0.  Nothing under oracle/ is linked into or
// called by the product library; only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may use it.
//
// PARITY PIN STATUS
//   * Pinned by the reference's own tests: alist parse/write (src/sparse.rs:548-647),
//     Phif64 flooding on the Johnson 4x6 code (src/decoder/flooding.rs:161-189),
//     encoder KATs (src/encoder.rs:128-197, src/encoder/staircase.rs:30-46),
//     puncturer KAT (src/simulation/puncturing.rs:118-129), BPSK KATs
//     (src/simulation/modulation.rs:294-309).  tests/test_oracle_kat.py runs them.
//   * "parity unpinned" for everything else: the reference has no test for the
//     horizontal-layered schedule, Tanh, Min*-approx, A-Min* or ANY i8 arithmetic,
//     and the reference itself cannot be run here (Rust edition 2024, no
//     cargo/rustc in the image, crates not vendored).  Those rules are restated
//     line by line from src/decoder/arithmetic.rs and cross-checked against an
//     independent pure-Python restatement (oracle/pyref.py) and the hand-derived
//     vectors of SURVEY.md §A.10.
//
// Every function cites the reference file:line it follows.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace oracle {

// ---------------------------------------------------------------- src/sparse.rs:23-26
struct SparseMatrix {
    std::vector<std::vector<size_t>> rows, cols;
    SparseMatrix() = default;
    SparseMatrix(size_t nrows, size_t ncols) : rows(nrows), cols(ncols) {}
    size_t num_rows() const { return rows.size(); }
    size_t num_cols() const { return cols.size(); }
    bool contains(size_t r, size_t c) const;          // src/sparse.rs:94-97
    void insert(size_t r, size_t c);                  // src/sparse.rs:114-119
    // src/sparse.rs:352-389.  Returns false and fills err on a parse error.
    static bool from_alist(const std::string& text, SparseMatrix* out, std::string* err);
    std::string alist(bool padding = true) const;     // src/sparse.rs:250-299
    size_t nnz() const;
};

struct DecodeResult {
    std::vector<uint8_t> codeword;
    size_t iterations = 0;
    bool success = false;
    bool error = false;   // the reference would have panicked (degree-0/1 check with min* rules)
};

// src/decoder.rs:19-35
class LdpcDecoder {
public:
    virtual ~LdpcDecoder() = default;
    virtual DecodeResult decode(const double* llrs, size_t n, size_t max_iterations) = 0;
    // posterior LLRs left by the last decode (converted to double); test hook.
    virtual std::vector<double> posteriors() const = 0;
    virtual void set_linear_search_send(bool) {}
    virtual size_t n() const = 0;
};

// src/decoder/factory.rs:240-277 — exact, case-sensitive names.
const std::vector<std::string>& implementation_names();
std::unique_ptr<LdpcDecoder> build_decoder(const std::string& implementation, const SparseMatrix& h);

// src/cli/ber.rs:219-229
bool parse_puncturing_pattern(const std::string& s, std::vector<bool>* out);

// src/simulation/puncturing.rs:11-110
struct Puncturer {
    std::vector<bool> pattern;
    size_t num_trues = 0;
    explicit Puncturer(const std::vector<bool>& p);
    template <class T> bool puncture(const std::vector<T>& cw, std::vector<T>* out) const;
    template <class T> bool depuncture(const T* llrs, size_t len, std::vector<T>* out) const;
    double rate() const { return double(pattern.size()) / double(num_trues); }
};

// src/encoder.rs:43-120
class Encoder {
public:
    static std::unique_ptr<Encoder> from_h(const SparseMatrix& h, std::string* err);
    // message: k bytes of 0/1 -> codeword n bytes of 0/1.
    void encode(const uint8_t* message, uint8_t* codeword) const;
    bool is_staircase() const { return staircase_; }
    size_t k() const { return k_; }
    size_t n() const { return n_; }
private:
    bool staircase_ = false;
    size_t k_ = 0, n_ = 0, m_ = 0;
    std::vector<std::vector<size_t>> h0_rows_;          // staircase: H0 rows
    std::vector<uint64_t> g0_;                          // dense: m x ceil(k/64) packed
    size_t g0_words_ = 0;
};

struct BerCounters {   // src/simulation/ber.rs:313-337, :498-581
    uint64_t num_frames = 0, bit_errors = 0, frame_errors = 0, false_decodes = 0;
    uint64_t total_iterations = 0, correct_iterations = 0;
    double elapsed_s = 0.0;
};

// src/simulation/ber.rs:297-368,:436-481 for one Eb/N0 point, BPSK/AWGN.
// Runs until `frames` frames are done (fixed work) or, if frames==0, until
// `max_frame_errors` frame errors were collected; nthreads workers.
bool ber_run(const SparseMatrix& h, const std::string& implementation, const std::string& puncturing,
             float ebn0_db, size_t max_iterations, uint64_t frames, uint64_t max_frame_errors,
             int nthreads, uint64_t seed, bool linear_search_send, BerCounters* out, std::string* err);

double noise_sigma(double rate, double bits_per_symbol, float ebn0_db);   // src/simulation/ber.rs:300-302

}  // namespace oracle
