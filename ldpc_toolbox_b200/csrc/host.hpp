// ldpc_toolbox_b200/csrc/host.hpp — host-side model of the reference's plugin boundary.
//
// Mirrors (same names, argument meaning and error behaviour), without sharing code with oracle/:
//   SparseMatrix::from_alist          reference src/sparse.rs:352-389
//   DecoderImplementation (36 names)  reference src/decoder/factory.rs:31-188,:240-277
//   LdpcDecoder / DecoderOutput       reference src/decoder.rs:19-48
//   DecoderFactory::build_decoder     reference src/decoder/factory.rs:19-25
//   Puncturer / pattern parser        reference src/simulation/puncturing.rs, src/cli/ber.rs:219-229
//   Encoder                           reference src/encoder.rs:43-120
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace ldpc {

// Parity-check matrix in the two orders the decoders depend on:
// row-major edges with each row in rows[r] order (= ascending column for alist input) and,
// per column, the list of those edge ids in cols[c] order (= file order of the column line).
struct Graph {
    int n = 0, m = 0, E = 0;
    std::vector<int> row_ptr;    // m+1
    std::vector<int> col_idx;    // E, variable of each row-major edge
    std::vector<int> col_ptr;    // n+1
    std::vector<int> col_edge;   // E, row-major edge id of each column-major slot
    std::vector<int> col_row;    // E, check of each column-major slot
    int max_row_deg = 0, min_row_deg = 0, max_col_deg = 0;
    // Parses alist text.  Returns false (and a message) where the reference returns Err; an
    // out-of-range row index, which makes the reference panic, is also reported as an error.
    static bool from_alist(const std::string& text, Graph* out, std::string* err);
    int k() const { return n - m; }
};

enum class Rule { Phi, Tanh, Minstarapprox, Aminstar };
enum class Dtype { F64, F32, I8 };
enum class Schedule { Flooding, HorizontalLayered };

struct DecoderImplementation {
    std::string name;
    Rule rule = Rule::Phi;
    Dtype dtype = Dtype::F64;
    Schedule schedule = Schedule::Flooding;
    bool jones = false, hardlimit = false, deg1clip = false;
    // FromStr of the reference (factory.rs:211-222): exact, case-sensitive.
    static bool parse(const std::string& s, DecoderImplementation* out);
    static const std::vector<std::string>& all_names();
};

bool parse_puncturing_pattern(const std::string& s, std::vector<bool>* out);

struct Puncturer {
    std::vector<bool> pattern;
    size_t num_trues = 0;
    explicit Puncturer(const std::vector<bool>& p);
    double rate() const { return double(pattern.size()) / double(num_trues); }
    // source index in the punctured vector for every codeword position, -1 if punctured.
    // false if punctured_len is not divisible by num_trues or does not expand to n_cw.
    bool depuncture_map(size_t punctured_len, size_t n_cw, std::vector<int>* map) const;
    // kept codeword positions in transmit order; false if n_cw % pattern.size() != 0
    bool puncture_map(size_t n_cw, std::vector<int>* kept) const;
};

// Systematic encoder (host).  Staircase codes keep H0 sparse; others get the dense, bit-packed
// G0 = H1^-1 H0 from a one-time GF(2) elimination, 64 columns per word.
struct EncoderPlan {
    bool staircase = false;
    int n = 0, m = 0, k = 0;
    std::vector<int> h0_ptr, h0_idx;       // CSR of H0 (m rows over k columns)
    std::vector<uint64_t> g0;              // m x words
    int words = 0;
    static bool from_graph(const Graph& g, EncoderPlan* out, std::string* err);
    void encode(const uint8_t* msg01, uint8_t* cw01) const;
};

}  // namespace ldpc
