// oracle/ldpc_oracle.cpp — TEST INFRASTRUCTURE ONLY (see ldpc_oracle.hpp header).
//
// C++17 restatement of the CPU algorithms of daniestevez/ldpc-toolbox v0.12.0
// for: alist parsing, the 36 decoder implementations (24 flooding + 12
// horizontal layered), the systematic encoder, the block puncturer and the
// BPSK/AWGN BER loop.  Written from the behaviour of the Rust sources; each
// block cites the file:line it follows.
#include "ldpc_oracle.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <functional>
#include <mutex>
#include <random>
#include <sstream>
#include <thread>

namespace oracle {

// =====================================================================================
// SparseMatrix  (src/sparse.rs)
// =====================================================================================
bool SparseMatrix::contains(size_t r, size_t c) const {
    const auto& col = cols[c];
    return std::find(col.begin(), col.end(), r) != col.end();
}

void SparseMatrix::insert(size_t r, size_t c) {
    if (!contains(r, c)) {
        rows[r].push_back(c);
        cols[c].push_back(r);
    }
}

size_t SparseMatrix::nnz() const {
    size_t e = 0;
    for (const auto& r : rows) e += r.size();
    return e;
}

namespace {
// Rust's str::split_whitespace + usize::from_str: tokens separated by ASCII/Unicode
// whitespace; a token is a number iff it is [+]?[0-9]+ and fits usize.
bool parse_usize(const std::string& tok, size_t* out) {
    size_t i = 0;
    if (tok.empty()) return false;
    if (tok[0] == '+') i = 1;
    if (i >= tok.size()) return false;
    unsigned long long v = 0;
    for (; i < tok.size(); ++i) {
        if (tok[i] < '0' || tok[i] > '9') return false;
        unsigned long long nv = v * 10ULL + (unsigned long long)(tok[i] - '0');
        if (nv < v) return false;
        v = nv;
    }
    *out = (size_t)v;
    return true;
}

std::vector<std::string> split_ws(const std::string& line) {
    std::vector<std::string> t;
    size_t i = 0;
    while (i < line.size()) {
        while (i < line.size() && std::isspace((unsigned char)line[i])) ++i;
        size_t j = i;
        while (j < line.size() && !std::isspace((unsigned char)line[j])) ++j;
        if (j > i) t.emplace_back(line.substr(i, j - i));
        i = j;
    }
    return t;
}
}  // namespace

// src/sparse.rs:352-389.  Only the first line and the column section are used;
// insertion order is file order, so rows[r] ends up ascending by column and
// cols[c] keeps the order of its line.
bool SparseMatrix::from_alist(const std::string& text, SparseMatrix* out, std::string* err) {
    std::vector<std::string> lines;
    {
        size_t pos = 0;
        while (true) {
            size_t nl = text.find('\n', pos);
            if (nl == std::string::npos) { lines.emplace_back(text.substr(pos)); break; }
            lines.emplace_back(text.substr(pos, nl - pos));
            pos = nl + 1;
        }
    }
    auto fail = [&](const char* m) { if (err) *err = m; return false; };
    if (lines.empty()) return fail("alist first line not found");
    auto sizes = split_ws(lines[0]);
    size_t ncols = 0, nrows = 0;
    if (sizes.size() < 1) return fail("alist first line does not contain enough elements");
    if (!parse_usize(sizes[0], &ncols)) return fail("ncols is not a number");
    if (sizes.size() < 2) return fail("alist first line does not contain enough elements");
    if (!parse_usize(sizes[1], &nrows)) return fail("nrows is not a number");
    SparseMatrix h(nrows, ncols);
    size_t li = 4;  // skip max weights, column weights, row weights
    for (size_t col = 0; col < ncols; ++col, ++li) {
        if (li >= lines.size()) return fail("alist does not contain expected number of lines");
        for (const auto& tok : split_ws(lines[li])) {
            size_t row;
            if (!parse_usize(tok, &row)) return fail("row value is not a number");
            if (row != 0) {
                // the reference indexes rows[row-1] and panics when out of range
                if (row - 1 >= nrows) return fail("row index out of range");
                h.insert(row - 1, col);
            }
        }
    }
    *out = std::move(h);
    return true;
}

// src/sparse.rs:250-299
std::string SparseMatrix::alist(bool padding) const {
    std::ostringstream w;
    w << num_cols() << " " << num_rows() << "\n";
    const std::vector<std::vector<size_t>>* dirs[2] = {&cols, &rows};
    size_t lens[2] = {0, 0};
    for (int d = 0; d < 2; ++d)
        for (const auto& el : *dirs[d]) lens[d] = std::max(lens[d], el.size());
    w << lens[0] << " " << lens[1] << "\n";
    for (int d = 0; d < 2; ++d) {
        bool first = true;
        for (const auto& el : *dirs[d]) {
            if (!first) w << " ";
            w << el.size();
            first = false;
        }
        w << "\n";
    }
    for (int d = 0; d < 2; ++d) {
        for (const auto& el : *dirs[d]) {
            std::vector<size_t> v(el);
            std::sort(v.begin(), v.end());
            for (size_t i = 0; i < v.size(); ++i) {
                if (i) w << " ";
                w << (v[i] + 1);
            }
            if (padding) {
                if (v.empty()) w << "0";
                size_t pad = lens[d] - std::max<size_t>(v.size(), 1);
                for (size_t i = 0; i < pad; ++i) w << " 0";
            }
            w << "\n";
        }
    }
    return w.str();
}

// =====================================================================================
// Decoder arithmetic  (src/decoder/arithmetic.rs)
//
// Each arithmetic exposes the same operations as the DecoderArithmetic trait
// (arithmetic.rs:44-137).  Exclusion is positional, which is identical to the
// reference's exclusion by node id because SparseMatrix::insert forbids
// duplicate entries (sparse.rs:114-119).  A `false` return means the reference
// would have panicked (`expect` on a degree-0/1 check node).
// =====================================================================================
namespace {

template <class F>
struct FloatBase {
    using Llr = F; using CheckMsg = F; using VarMsg = F; using VarLlr = F;
    static F quantize(double llr) { return (F)llr; }               // :194-196 `llr as $f`
    static bool hard(F llr) { return llr <= F(0); }                // :198-200
    static F llr_to_var_message(F l) { return l; }
    static F llr_to_var_llr(F l) { return l; }
    static F var_llr_to_llr(F l) { return l; }
    static double to_double(F l) { return (double)l; }
    // send_var_messages_no_clip, arithmetic.rs:140-156.  The check-message sum is
    // accumulated on its own (iterator .sum()) and then added to the input LLR.
    static F send_var(F input, const F* in, size_t d, F* out) {
        F sum = F(0);
        for (size_t i = 0; i < d; ++i) sum += in[i];
        F llr = input + sum;
        for (size_t i = 0; i < d; ++i) out[i] = llr - in[i];
        return llr;
    }
};

inline float fmax_(float a, float b) { return std::fmax(a, b); }
inline double fmax_(double a, double b) { return std::fmax(a, b); }
inline float fmin_(float a, float b) { return std::fmin(a, b); }
inline double fmin_(double a, double b) { return std::fmin(a, b); }

// ---- Phi, arithmetic.rs:158-298
template <class F>
struct PhiArith : FloatBase<F> {
    std::vector<F> phis;
    static F phi(F x) {                                              // :180-185
        x = fmax_(x, F(1e-30));
        return -std::log(std::tanh(F(0.5) * x));
    }
    bool send_check(const F* in, size_t d, F* out) {                 // :214-246
        unsigned sign = 0;
        F sum = F(0);
        if (phis.size() < d) phis.resize(d, F(0));
        for (size_t i = 0; i < d; ++i) {
            F x = in[i];
            F p = phi(std::fabs(x));
            phis[i] = p;
            sum += p;
            if (x < F(0)) sign ^= 1;
        }
        for (size_t i = 0; i < d; ++i) {
            F x = in[i];
            F y = phi(sum - phis[i]);
            unsigned s = (x < F(0)) ? (sign ^ 1u) : sign;
            out[i] = (s == 0) ? y : -y;
        }
        return true;
    }
    bool update_row(F* rcv, const size_t* dest, size_t d, F* vars) { // :260-292
        unsigned sign = 0;
        F sum = F(0);
        if (phis.size() < d) phis.resize(d, F(0));
        for (size_t i = 0; i < d; ++i) {
            F x = vars[dest[i]] - rcv[i];
            F p = phi(std::fabs(x));
            phis[i] = p;
            sum += p;
            if (x < F(0)) sign ^= 1;
        }
        for (size_t i = 0; i < d; ++i) {
            F x = vars[dest[i]] - rcv[i];
            F r = phi(sum - phis[i]);
            unsigned s = (x < F(0)) ? (sign ^ 1u) : sign;
            r = (s == 0) ? r : -r;
            rcv[i] = r;
            vars[dest[i]] = x + r;
        }
        return true;
    }
};

// ---- Tanh, arithmetic.rs:300-435
template <class F>
struct TanhArith : FloatBase<F> {
    std::vector<F> tanhs;
    static constexpr F clampv() { return sizeof(F) == 8 ? F(18.0) : F(9.0); }   // :433-435
    static F clamp(F x) {  // Rust clamp: NaN stays NaN
        const F c = clampv();
        if (x < -c) return -c;
        if (x > c) return c;
        return x;
    }
    bool send_check(const F* in, size_t d, F* out) {                 // :347-379
        if (tanhs.size() < d) tanhs.resize(d, F(0));
        for (size_t i = 0; i < d; ++i) tanhs[i] = std::tanh(clamp(F(0.5) * in[i]));
        for (size_t j = 0; j < d; ++j) {
            F product = F(1);
            for (size_t i = 0; i < d; ++i)
                if (i != j) product *= tanhs[i];
            out[j] = F(2) * std::atanh(product);
        }
        return true;
    }
    bool update_row(F* rcv, const size_t* dest, size_t d, F* vars) { // :393-426
        if (tanhs.size() < d) tanhs.resize(d, F(0));
        for (size_t i = 0; i < d; ++i) tanhs[i] = std::tanh(clamp(F(0.5) * (vars[dest[i]] - rcv[i])));
        for (size_t j = 0; j < d; ++j) {
            F product = F(1);
            for (size_t i = 0; i < d; ++i)
                if (i != j) product *= tanhs[i];
            F r = F(2) * std::atanh(product);
            vars[dest[j]] += r - rcv[j];
            rcv[j] = r;
        }
        return true;
    }
};

// ---- Min*-approx float, arithmetic.rs:437-580
template <class F>
struct MinstarApproxFArith : FloatBase<F> {
    std::vector<F> minstars;
    static F g(F x, F y) {                                            // :510
        return fmax_(fmin_(x, y) - std::log1p(std::exp(-std::fabs(x - y))), F(0));
    }
    // fold over all i != j in order; false if there is no other input
    template <class Get>
    static bool fold_excluding(size_t d, size_t j, Get get, F* outv) {
        unsigned sign = 0;
        bool have = false;
        F acc = F(0);
        for (size_t i = 0; i < d; ++i) {
            if (i == j) continue;
            F x = get(i);
            if (x < F(0)) sign ^= 1;
            x = std::fabs(x);
            acc = have ? g(x, acc) : x;
            have = true;
        }
        if (!have) return false;                                      // :513-514 expect()
        *outv = (sign == 0) ? acc : -acc;
        return true;
    }
    bool send_check(const F* in, size_t d, F* out) {                 // :487-521
        for (size_t j = 0; j < d; ++j)
            if (!fold_excluding(d, j, [&](size_t i) { return in[i]; }, &out[j])) return false;
        return true;
    }
    bool update_row(F* rcv, const size_t* dest, size_t d, F* vars) { // :535-574
        if (minstars.size() < d) minstars.resize(d, F(0));
        for (size_t j = 0; j < d; ++j)
            if (!fold_excluding(d, j, [&](size_t i) { return vars[dest[i]] - rcv[i]; }, &minstars[j])) return false;
        for (size_t i = 0; i < d; ++i) {
            vars[dest[i]] += minstars[i] - rcv[i];
            rcv[i] = minstars[i];
        }
        return true;
    }
};

// ---- A-Min* float, arithmetic.rs:899-1072
template <class F>
struct AminstarFArith : FloatBase<F> {
    static F h(F x, F y) {                                            // :965-966
        return fmin_(x, y) - std::log1p(std::exp(-std::fabs(x - y))) + std::log1p(std::exp(-(x + y)));
    }
    // returns false on empty / single input (reference panics)
    template <class Get>
    static bool core(size_t d, Get get, size_t* argmin_out, F* delta_min_edge, F* delta_others, unsigned* sign_out) {
        if (d == 0) return false;                                     // :952 expect
        size_t argmin = 0;                                            // first minimum (min_by keeps the first)
        F best = std::fabs(get(0));
        for (size_t i = 1; i < d; ++i) {
            F a = std::fabs(get(i));
            if (a < best) { best = a; argmin = i; }
        }
        unsigned sign = 0;
        bool have = false;
        F delta = F(0);
        for (size_t j = 0; j < d; ++j) {
            F x = get(j);
            if (x < F(0)) sign ^= 1;
            if (j != argmin) {
                x = std::fabs(x);
                delta = have ? h(x, delta) : x;
                have = true;
            }
        }
        if (!have) return false;                                      // :971 expect
        *delta_min_edge = delta;
        F vmin = std::fabs(get(argmin));
        *delta_others = h(delta, vmin);                               // :982-984
        *argmin_out = argmin;
        *sign_out = sign;
        return true;
    }
    bool send_check(const F* in, size_t d, F* out) {                 // :942-999
        size_t argmin; F dmin, doth; unsigned sign;
        if (!core(d, [&](size_t i) { return in[i]; }, &argmin, &dmin, &doth, &sign)) return false;
        for (size_t i = 0; i < d; ++i) {
            F mag = (i == argmin) ? dmin : doth;
            bool neg = (sign != 0) ^ (in[i] < F(0));
            out[i] = neg ? -mag : mag;
        }
        return true;
    }
    bool update_row(F* rcv, const size_t* dest, size_t d, F* vars) { // :1013-1066
        size_t argmin; F dmin, doth; unsigned sign;
        if (!core(d, [&](size_t i) { return vars[dest[i]] - rcv[i]; }, &argmin, &dmin, &doth, &sign)) return false;
        F xmin = vars[dest[argmin]] - rcv[argmin];
        F msgmin_rcv = ((sign != 0) ^ (xmin < F(0))) ? -dmin : dmin;
        for (size_t j = 0; j < d; ++j) {
            F x = vars[dest[j]] - rcv[j];
            F r;
            if (j == argmin) r = msgmin_rcv;
            else r = ((sign != 0) ^ (x < F(0))) ? -doth : doth;
            rcv[j] = r;
            vars[dest[j]] = x + r;
        }
        return true;
    }
};

// ---- 8-bit quantised arithmetics, arithmetic.rs:582-897 and :1074-1304
struct I8Table {
    std::vector<int8_t> table;
    I8Table() {                                                       // :588-597
        for (int t = 0; t <= 127; ++t) {
            double x = std::round(8.0 * std::log1p(std::exp(-((double)t / 8.0))));
            int8_t xi = (int8_t)x;
            if (xi > 0) table.push_back(xi); else break;
        }
    }
    int lookup(int x) const { return (x >= 0 && (size_t)x < table.size()) ? table[(size_t)x] : 0; }  // :604-607
};

inline int8_t clip_i8(int x) { return x >= 127 ? 127 : (x <= -127 ? -127 : (int8_t)x); }   // :609-617

template <bool JONES, bool HARDLIMIT, bool DEG1CLIP>
struct I8Base {
    using Llr = int8_t; using CheckMsg = int8_t; using VarMsg = int8_t; using VarLlr = int16_t;
    I8Table T;
    static int8_t quantize(double llr) {                              // :690-699
        double x = 8.0 * llr;
        if (x >= 127.0) return 127;
        if (x <= -127.0) return -127;
        double r = std::round(x);       // f64::round: half away from zero
        if (std::isnan(r)) return 0;    // `as i8` maps NaN to 0
        return (int8_t)r;
    }
    static bool hard(int8_t l) { return l <= 0; }                     // :701-703
    static int8_t llr_to_var_message(int8_t l) { return l; }
    static int16_t llr_to_var_llr(int8_t l) { return (int16_t)l; }
    static int8_t var_llr_to_llr(int16_t v) { return clip_i8(v); }    // :713-715
    static double to_double(int8_t l) { return (double)l; }
    static int hardlimit(int x) {                                     // :812-824
        if (!HARDLIMIT) return x;
        if (x <= -100) return -127;
        if (x >= 100) return 127;
        return x;
    }
    // impl_send_var_messages_i8, :622-654
    static int8_t send_var(int8_t input, const int8_t* in, size_t d, int8_t* out) {
        bool degree_one = (d == 1);
        int inp = input;
        if (DEG1CLIP && degree_one) {                                 // :826-842
            if (inp <= -116) inp = -116; else if (inp >= 116) inp = 116;
        }
        int sum = 0;
        for (size_t i = 0; i < d; ++i) sum += in[i];
        int llr = (int16_t)(inp + sum);
        if (JONES) llr = clip_i8(llr);                                // :806-810
        for (size_t i = 0; i < d; ++i) out[i] = clip_i8(llr - in[i]);
        return clip_i8(llr);
    }
};

template <bool JONES, bool HARDLIMIT, bool DEG1CLIP>
struct MinstarApproxI8Arith : I8Base<JONES, HARDLIMIT, DEG1CLIP> {
    using B = I8Base<JONES, HARDLIMIT, DEG1CLIP>;
    std::vector<int8_t> minstars;
    int g(int x, int y) const {                                       // :741
        int v = std::min(x, y) - this->T.lookup(std::abs(x - y));
        return v > 0 ? v : 0;
    }
    template <class Get>
    bool fold_excluding(size_t d, size_t j, Get get, int8_t* outv) const {
        unsigned sign = 0;
        bool have = false;
        int acc = 0;
        for (size_t i = 0; i < d; ++i) {
            if (i == j) continue;
            int x = get(i);
            if (x < 0) sign ^= 1;
            x = std::abs(x);
            acc = have ? g(x, acc) : x;
            have = true;
        }
        if (!have) return false;                                      // :744-745 expect
        int m = (sign == 0) ? acc : -acc;
        *outv = (int8_t)B::hardlimit(m);
        return true;
    }
    bool send_check(const int8_t* in, size_t d, int8_t* out) {        // :718-754
        for (size_t j = 0; j < d; ++j)
            if (!fold_excluding(d, j, [&](size_t i) { return (int)in[i]; }, &out[j])) return false;
        return true;
    }
    bool update_row(int8_t* rcv, const size_t* dest, size_t d, int16_t* vars) {   // :759-801
        if (minstars.size() < d) minstars.resize(d, 0);
        for (size_t j = 0; j < d; ++j)
            if (!fold_excluding(d, j, [&](size_t i) { return (int)clip_i8((int)vars[dest[i]] - (int)rcv[i]); }, &minstars[j]))
                return false;
        for (size_t i = 0; i < d; ++i) {
            vars[dest[i]] = (int16_t)(vars[dest[i]] + ((int)minstars[i] - (int)rcv[i]));
            rcv[i] = minstars[i];
        }
        return true;
    }
};

template <bool JONES, bool HARDLIMIT, bool DEG1CLIP>
struct AminstarI8Arith : I8Base<JONES, HARDLIMIT, DEG1CLIP> {
    using B = I8Base<JONES, HARDLIMIT, DEG1CLIP>;
    int h(int x, int y) const {                                       // :1155-1157
        int sat = std::min(x + y, 127);                               // i8 saturating_add
        int v = std::min(x, y) - this->T.lookup(std::abs(x - y)) + this->T.lookup(sat);
        return v > 0 ? v : 0;
    }
    template <class Get>
    bool core(size_t d, Get get, size_t* argmin_out, int* dmin_hl, int* doth_hl, unsigned* sign_out) const {
        if (d == 0) return false;
        size_t argmin = 0;                                            // min_by_key keeps the first minimum
        int best = std::abs(get(0));
        for (size_t i = 1; i < d; ++i) {
            int a = std::abs(get(i));
            if (a < best) { best = a; argmin = i; }
        }
        unsigned sign = 0;
        bool have = false;
        int delta = 0;
        for (size_t j = 0; j < d; ++j) {
            int x = get(j);
            if (x < 0) sign ^= 1;
            if (j != argmin) {
                x = std::abs(x);
                delta = have ? h(x, delta) : x;
                have = true;
            }
        }
        if (!have) return false;
        *dmin_hl = B::hardlimit(delta);                               // :1162
        int vmin = std::abs(get(argmin));
        int d2 = h(delta, vmin);                                      // :1173-1176
        *doth_hl = B::hardlimit(d2);
        *argmin_out = argmin;
        *sign_out = sign;
        return true;
    }
    bool send_check(const int8_t* in, size_t d, int8_t* out) {        // :1130-1192
        size_t argmin; int dmin, doth; unsigned sign;
        if (!core(d, [&](size_t i) { return (int)in[i]; }, &argmin, &dmin, &doth, &sign)) return false;
        for (size_t i = 0; i < d; ++i) {
            int mag = (i == argmin) ? dmin : doth;
            bool neg = (sign != 0) ^ (in[i] < 0);
            out[i] = (int8_t)(neg ? -mag : mag);
        }
        return true;
    }
    bool update_row(int8_t* rcv, const size_t* dest, size_t d, int16_t* vars) {   // :1197-1257
        size_t argmin; int dmin, doth; unsigned sign;
        auto get = [&](size_t i) { return (int)clip_i8((int)vars[dest[i]] - (int)rcv[i]); };
        if (!core(d, get, &argmin, &dmin, &doth, &sign)) return false;
        int msgmin = get(argmin);
        int msgmin_rcv = ((sign != 0) ^ (msgmin < 0)) ? -dmin : dmin;
        for (size_t j = 0; j < d; ++j) {
            int x = (int)vars[dest[j]] - (int)rcv[j];                 // unclipped difference, :1244
            int r;
            if (j == argmin) r = msgmin_rcv;
            else r = ((sign != 0) ^ (x < 0)) ? -doth : doth;
            vars[dest[j]] = (int16_t)(x + r);
            rcv[j] = (int8_t)r;
        }
        return true;
    }
};

// =====================================================================================
// src/decoder.rs:157-174
// =====================================================================================
template <class T, class HD>
bool check_llrs(const SparseMatrix& h, const T* llrs, HD hd) {
    for (size_t r = 0; r < h.num_rows(); ++r) {
        unsigned cnt = 0;
        for (size_t c : h.rows[r]) cnt += hd(llrs[c]) ? 1u : 0u;
        if (cnt % 2 == 1) return false;
    }
    return true;
}

template <class T, class HD>
std::vector<uint8_t> hard_decisions(const T* llrs, size_t n, HD hd) {
    std::vector<uint8_t> v(n);
    for (size_t i = 0; i < n; ++i) v[i] = hd(llrs[i]) ? 1 : 0;
    return v;
}

// =====================================================================================
// Flooding decoder, src/decoder/flooding.rs:13-125.
//
// Message storage mirrors decoder.rs:85-155: variable->check messages are kept per
// check in rows[c] order, check->variable messages per variable in cols[v] order.
// `send` locates the slot of (source -> destination); the reference does a linear
// search (decoder.rs:111-117).  Here the slot index is precomputed; setting
// linear_search=true performs the reference's search instead (same result, used only
// to time the reference's actual cost in the CPU baseline).
// =====================================================================================
template <class A>
class FloodingDecoder final : public LdpcDecoder {
public:
    FloodingDecoder(const SparseMatrix& h) : h_(h) {
        size_t n = h.num_cols(), m = h.num_rows();
        input_.assign(n, typename A::Llr());
        output_.assign(n, typename A::Llr());
        row_ptr_.assign(m + 1, 0);
        col_ptr_.assign(n + 1, 0);
        for (size_t r = 0; r < m; ++r) row_ptr_[r + 1] = row_ptr_[r] + h.rows[r].size();
        for (size_t c = 0; c < n; ++c) col_ptr_[c + 1] = col_ptr_[c] + h.cols[c].size();
        size_t e = row_ptr_[m];
        v2c_.assign(e, typename A::VarMsg());
        c2v_.assign(e, typename A::CheckMsg());
        row2col_.resize(e);
        col2row_.resize(e);
        for (size_t r = 0; r < m; ++r)
            for (size_t j = 0; j < h.rows[r].size(); ++j) {
                size_t v = h.rows[r][j];
                const auto& col = h.cols[v];
                size_t p = std::find(col.begin(), col.end(), r) - col.begin();
                row2col_[row_ptr_[r] + j] = col_ptr_[v] + p;
                col2row_[col_ptr_[v] + p] = row_ptr_[r] + j;
            }
        size_t maxd = 0;
        for (const auto& r : h.rows) maxd = std::max(maxd, r.size());
        for (const auto& c : h.cols) maxd = std::max(maxd, c.size());
        tmp_c_.resize(maxd);
        tmp_v_.resize(maxd);
    }
    size_t n() const override { return h_.num_cols(); }
    void set_linear_search_send(bool b) override { linear_ = b; }
    std::vector<double> posteriors() const override {
        std::vector<double> p(output_.size());
        for (size_t i = 0; i < p.size(); ++i) p[i] = A::to_double(output_[i]);
        return p;
    }

    DecodeResult decode(const double* llrs, size_t n, size_t max_iterations) override {   // :51-86
        DecodeResult res;
        if (n != input_.size()) { res.error = true; return res; }    // :56 assert_eq!
        auto raw_hd = [](double x) { return x <= 0.0; };
        if (check_llrs(h_, llrs, raw_hd)) {                          // :57-64
            res.codeword = hard_decisions(llrs, n, raw_hd);
            res.iterations = 0;
            res.success = true;
            return res;
        }
        initialize(llrs);
        for (size_t it = 1; it <= max_iterations; ++it) {
            if (!process_check_nodes()) { res.error = true; return res; }
            process_variable_nodes();
            if (check_llrs(h_, output_.data(), A::hard)) {
                res.codeword = hard_decisions(output_.data(), n, A::hard);
                res.iterations = it;
                res.success = true;
                return res;
            }
        }
        res.codeword = hard_decisions(output_.data(), n, A::hard);   // :81-85
        res.iterations = max_iterations;
        res.success = false;
        return res;
    }

private:
    size_t find_slot_in_check(size_t c, size_t v) const {            // decoder.rs:111-117
        const auto& row = h_.rows[c];
        for (size_t j = 0; j < row.size(); ++j) if (row[j] == v) return row_ptr_[c] + j;
        return (size_t)-1;
    }
    size_t find_slot_in_var(size_t v, size_t c) const {
        const auto& col = h_.cols[v];
        for (size_t j = 0; j < col.size(); ++j) if (col[j] == c) return col_ptr_[v] + j;
        return (size_t)-1;
    }
    void initialize(const double* llrs) {                            // :88-100
        for (size_t v = 0; v < input_.size(); ++v) input_[v] = A::quantize(llrs[v]);
        for (size_t v = 0; v < input_.size(); ++v) {
            auto msg = A::llr_to_var_message(input_[v]);
            for (size_t p = col_ptr_[v]; p < col_ptr_[v + 1]; ++p) {
                size_t slot = linear_ ? find_slot_in_check(h_.cols[v][p - col_ptr_[v]], v) : col2row_[p];
                v2c_[slot] = msg;
            }
        }
    }
    bool process_check_nodes() {                                     // :102-109
        for (size_t c = 0; c < h_.num_rows(); ++c) {
            size_t b = row_ptr_[c], d = row_ptr_[c + 1] - b;
            if (!arith_.send_check(v2c_.data() + b, d, tmp_c_.data())) return false;   // reference panics
            for (size_t j = 0; j < d; ++j) {
                size_t slot = linear_ ? find_slot_in_var(h_.rows[c][j], c) : row2col_[b + j];
                c2v_[slot] = tmp_c_[j];
            }
        }
        return true;
    }
    void process_variable_nodes() {                                  // :111-125
        for (size_t v = 0; v < h_.num_cols(); ++v) {
            size_t b = col_ptr_[v], d = col_ptr_[v + 1] - b;
            output_[v] = A::send_var(input_[v], c2v_.data() + b, d, tmp_v_.data());
            for (size_t j = 0; j < d; ++j) {
                size_t slot = linear_ ? find_slot_in_check(h_.cols[v][j], v) : col2row_[b + j];
                v2c_[slot] = tmp_v_[j];
            }
        }
    }
    SparseMatrix h_;
    A arith_;
    bool linear_ = false;
    std::vector<typename A::Llr> input_, output_;
    std::vector<typename A::VarMsg> v2c_, tmp_v_;
    std::vector<typename A::CheckMsg> c2v_, tmp_c_;
    std::vector<size_t> row_ptr_, col_ptr_, row2col_, col2row_;
};

// =====================================================================================
// Horizontal layered decoder, src/decoder/horizontal_layered.rs:17-110
// =====================================================================================
template <class A>
class LayeredDecoder final : public LdpcDecoder {
public:
    LayeredDecoder(const SparseMatrix& h) : h_(h) {
        size_t m = h.num_rows();
        llrs_.assign(h.num_cols(), typename A::VarLlr());
        row_ptr_.assign(m + 1, 0);
        for (size_t r = 0; r < m; ++r) row_ptr_[r + 1] = row_ptr_[r] + h.rows[r].size();
        rcv_.assign(row_ptr_[m], typename A::CheckMsg());
        dest_.reserve(row_ptr_[m]);
        for (size_t r = 0; r < m; ++r) for (size_t c : h.rows[r]) dest_.push_back(c);
    }
    size_t n() const override { return h_.num_cols(); }
    std::vector<double> posteriors() const override {
        std::vector<double> p(llrs_.size());
        for (size_t i = 0; i < p.size(); ++i) p[i] = (double)llrs_[i];
        return p;
    }
    DecodeResult decode(const double* llrs, size_t n, size_t max_iterations) override {   // :49-88
        DecodeResult res;
        if (n != llrs_.size()) { res.error = true; return res; }
        auto raw_hd = [](double x) { return x <= 0.0; };
        if (check_llrs(h_, llrs, raw_hd)) {
            res.codeword = hard_decisions(llrs, n, raw_hd);
            res.iterations = 0;
            res.success = true;
            return res;
        }
        for (size_t v = 0; v < n; ++v) llrs_[v] = A::llr_to_var_llr(A::quantize(llrs[v]));   // :90-96
        std::fill(rcv_.begin(), rcv_.end(), typename A::CheckMsg());                          // :97-102
        auto hd = [](typename A::VarLlr x) { return A::hard(A::var_llr_to_llr(x)); };
        for (size_t it = 1; it <= max_iterations; ++it) {
            for (size_t r = 0; r < h_.num_rows(); ++r) {                                      // :105-110
                size_t b = row_ptr_[r], d = row_ptr_[r + 1] - b;
                if (!arith_.update_row(rcv_.data() + b, dest_.data() + b, d, llrs_.data())) {
                    res.error = true;   // the reference panics (expect on a degree-0/1 row)
                    return res;
                }
            }
            if (check_llrs(h_, llrs_.data(), hd)) {
                res.codeword = hard_decisions(llrs_.data(), n, hd);
                res.iterations = it;
                res.success = true;
                return res;
            }
        }
        res.codeword = hard_decisions(llrs_.data(), n, hd);
        res.iterations = max_iterations;
        res.success = false;
        return res;
    }
private:
    SparseMatrix h_;
    A arith_;
    std::vector<typename A::VarLlr> llrs_;
    std::vector<typename A::CheckMsg> rcv_;
    std::vector<size_t> row_ptr_, dest_;
};

}  // namespace

// =====================================================================================
// Factory, src/decoder/factory.rs:240-277
// =====================================================================================
namespace {
struct ImplEntry {
    const char* name;
    std::function<std::unique_ptr<LdpcDecoder>(const SparseMatrix&)> make;
};

template <class A> std::unique_ptr<LdpcDecoder> mk_flood(const SparseMatrix& h) { return std::make_unique<FloodingDecoder<A>>(h); }
template <class A> std::unique_ptr<LdpcDecoder> mk_hl(const SparseMatrix& h) { return std::make_unique<LayeredDecoder<A>>(h); }

const std::vector<ImplEntry>& impl_table() {
    static const std::vector<ImplEntry> t = {
        {"Phif64", mk_flood<PhiArith<double>>},
        {"Phif32", mk_flood<PhiArith<float>>},
        {"Tanhf64", mk_flood<TanhArith<double>>},
        {"Tanhf32", mk_flood<TanhArith<float>>},
        {"Minstarapproxf64", mk_flood<MinstarApproxFArith<double>>},
        {"Minstarapproxf32", mk_flood<MinstarApproxFArith<float>>},
        {"Minstarapproxi8", mk_flood<MinstarApproxI8Arith<false, false, false>>},
        {"Minstarapproxi8Jones", mk_flood<MinstarApproxI8Arith<true, false, false>>},
        {"Minstarapproxi8PartialHardLimit", mk_flood<MinstarApproxI8Arith<false, true, false>>},
        {"Minstarapproxi8JonesPartialHardLimit", mk_flood<MinstarApproxI8Arith<true, true, false>>},
        {"Minstarapproxi8Deg1Clip", mk_flood<MinstarApproxI8Arith<false, false, true>>},
        {"Minstarapproxi8JonesDeg1Clip", mk_flood<MinstarApproxI8Arith<true, false, true>>},
        {"Minstarapproxi8PartialHardLimitDeg1Clip", mk_flood<MinstarApproxI8Arith<false, true, true>>},
        {"Minstarapproxi8JonesPartialHardLimitDeg1Clip", mk_flood<MinstarApproxI8Arith<true, true, true>>},
        {"Aminstarf64", mk_flood<AminstarFArith<double>>},
        {"Aminstarf32", mk_flood<AminstarFArith<float>>},
        {"Aminstari8", mk_flood<AminstarI8Arith<false, false, false>>},
        {"Aminstari8Jones", mk_flood<AminstarI8Arith<true, false, false>>},
        {"Aminstari8PartialHardLimit", mk_flood<AminstarI8Arith<false, true, false>>},
        {"Aminstari8JonesPartialHardLimit", mk_flood<AminstarI8Arith<true, true, false>>},
        {"Aminstari8Deg1Clip", mk_flood<AminstarI8Arith<false, false, true>>},
        {"Aminstari8JonesDeg1Clip", mk_flood<AminstarI8Arith<true, false, true>>},
        {"Aminstari8PartialHardLimitDeg1Clip", mk_flood<AminstarI8Arith<false, true, true>>},
        {"Aminstari8JonesPartialHardLimitDeg1Clip", mk_flood<AminstarI8Arith<true, true, true>>},
        {"HLPhif64", mk_hl<PhiArith<double>>},
        {"HLPhif32", mk_hl<PhiArith<float>>},
        {"HLTanhf64", mk_hl<TanhArith<double>>},
        {"HLTanhf32", mk_hl<TanhArith<float>>},
        {"HLMinstarapproxf64", mk_hl<MinstarApproxFArith<double>>},
        {"HLMinstarapproxf32", mk_hl<MinstarApproxFArith<float>>},
        {"HLMinstarapproxi8", mk_hl<MinstarApproxI8Arith<false, false, false>>},
        {"HLMinstarapproxi8PartialHardLimit", mk_hl<MinstarApproxI8Arith<false, true, false>>},
        {"HLAminstarf64", mk_hl<AminstarFArith<double>>},
        {"HLAminstarf32", mk_hl<AminstarFArith<float>>},
        {"HLAminstari8", mk_hl<AminstarI8Arith<false, false, false>>},
        {"HLAminstari8PartialHardLimit", mk_hl<AminstarI8Arith<false, true, false>>},
    };
    return t;
}
}  // namespace

const std::vector<std::string>& implementation_names() {
    static const std::vector<std::string> names = [] {
        std::vector<std::string> v;
        for (const auto& e : impl_table()) v.emplace_back(e.name);
        return v;
    }();
    return names;
}

std::unique_ptr<LdpcDecoder> build_decoder(const std::string& implementation, const SparseMatrix& h) {
    for (const auto& e : impl_table())
        if (implementation == e.name) return e.make(h);
    return nullptr;   // "invalid decoder implementation", factory.rs:219
}

// =====================================================================================
// Puncturing, src/cli/ber.rs:219-229 and src/simulation/puncturing.rs
// =====================================================================================
bool parse_puncturing_pattern(const std::string& s, std::vector<bool>* out) {
    out->clear();
    size_t pos = 0;
    while (true) {
        size_t c = s.find(',', pos);
        std::string a = s.substr(pos, c == std::string::npos ? std::string::npos : c - pos);
        if (a == "0") out->push_back(false);
        else if (a == "1") out->push_back(true);
        else return false;
        if (c == std::string::npos) break;
        pos = c + 1;
    }
    return true;
}

Puncturer::Puncturer(const std::vector<bool>& p) : pattern(p) {
    for (bool b : p) num_trues += b ? 1 : 0;
}

template <class T>
bool Puncturer::puncture(const std::vector<T>& cw, std::vector<T>* out) const {   // puncturing.rs:47-75
    size_t plen = pattern.size();
    if (cw.size() % plen != 0) return false;
    size_t bs = cw.size() / plen;
    out->assign(bs * num_trues, T());
    size_t j = 0;
    for (size_t k = 0; k < plen; ++k) {
        if (!pattern[k]) continue;
        std::copy(cw.begin() + k * bs, cw.begin() + (k + 1) * bs, out->begin() + j * bs);
        ++j;
    }
    return true;
}

template <class T>
bool Puncturer::depuncture(const T* llrs, size_t len, std::vector<T>* out) const {  // puncturing.rs:85-101
    if (num_trues == 0 || len % num_trues != 0) return false;
    size_t bs = len / num_trues;
    out->assign(pattern.size() * bs, T());
    size_t j = 0;
    for (size_t k = 0; k < pattern.size(); ++k) {
        if (!pattern[k]) continue;
        std::copy(llrs + j * bs, llrs + (j + 1) * bs, out->begin() + k * bs);
        ++j;
    }
    return true;
}
template bool Puncturer::puncture<uint8_t>(const std::vector<uint8_t>&, std::vector<uint8_t>*) const;
template bool Puncturer::puncture<double>(const std::vector<double>&, std::vector<double>*) const;
template bool Puncturer::depuncture<double>(const double*, size_t, std::vector<double>*) const;
template bool Puncturer::depuncture<float>(const float*, size_t, std::vector<float>*) const;

// =====================================================================================
// Encoder, src/encoder.rs:59-120, src/encoder/staircase.rs:3-24, src/linalg.rs:8-66
// =====================================================================================
namespace {
bool h_is_staircase(const SparseMatrix& h) {                            // staircase.rs:3-24
    size_t n = h.num_rows(), m = h.num_cols();
    size_t num_checked = 0;
    for (size_t j = 0; j < n; ++j)
        for (size_t k : h.rows[j]) {
            if (k >= m - n) {
                if (j == 0 && k != m - n) return false;
                if (j != 0 && k != m - n + j - 1 && k != m - n + j) return false;
                ++num_checked;
            }
        }
    return num_checked == 2 * n - 1;
}
}  // namespace

std::unique_ptr<Encoder> Encoder::from_h(const SparseMatrix& h, std::string* err) {
    size_t nr = h.num_rows(), nc = h.num_cols();
    if (nr > nc || nr == 0) {   // the reference underflows `m - n` / builds an empty code
        if (nr > nc) { if (err) *err = "more rows than columns"; return nullptr; }
    }
    auto enc = std::unique_ptr<Encoder>(new Encoder());
    enc->n_ = nc; enc->m_ = nr; enc->k_ = nc - nr;
    size_t k = enc->k_;
    if (nr > 0 && h_is_staircase(h)) {                                  // encoder.rs:63-74
        enc->staircase_ = true;
        enc->h0_rows_.resize(nr);
        for (size_t j = 0; j < nr; ++j)
            for (size_t c : h.rows[j]) if (c < k) enc->h0_rows_[j].push_back(c);
        return enc;
    }
    // encoder.rs:76-93: A = [H1 H0], Gauss-Jordan to [I | H1^-1 H0]; bit-packed here,
    // 64 columns per word (the result G0 = H1^-1 H0 is unique, so the pivoting order of
    // linalg.rs:8-66 does not matter).
    size_t words = (nc + 63) / 64;
    std::vector<uint64_t> a(nr * words, 0);
    for (size_t j = 0; j < nr; ++j)
        for (size_t c : h.rows[j]) {
            size_t t = c < k ? c + nr : c - k;
            a[j * words + t / 64] ^= (uint64_t)1 << (t % 64);   // insert() never duplicates
        }
    auto bit = [&](size_t r, size_t c) { return (a[r * words + c / 64] >> (c % 64)) & 1; };
    for (size_t j = 0; j < nr; ++j) {
        size_t p = j;
        while (p < nr && !bit(p, j)) ++p;
        if (p == nr) { if (err) *err = "the square matrix formed by the last columns of the parity check is not invertible"; return nullptr; }
        if (p != j) for (size_t w = 0; w < words; ++w) std::swap(a[j * words + w], a[p * words + w]);
        for (size_t t = 0; t < nr; ++t) {
            if (t != j && bit(t, j))
                for (size_t w = j / 64; w < words; ++w) a[t * words + w] ^= a[j * words + w];
        }
    }
    enc->g0_words_ = (k + 63) / 64;
    enc->g0_.assign(nr * enc->g0_words_, 0);
    for (size_t j = 0; j < nr; ++j)
        for (size_t c = 0; c < k; ++c)
            if (bit(j, nr + c)) enc->g0_[j * enc->g0_words_ + c / 64] |= (uint64_t)1 << (c % 64);
    return enc;
}

void Encoder::encode(const uint8_t* message, uint8_t* codeword) const {  // encoder.rs:99-120
    std::memcpy(codeword, message, k_);
    uint8_t* parity = codeword + k_;
    if (staircase_) {
        for (size_t j = 0; j < m_; ++j) {
            uint8_t s = 0;
            for (size_t c : h0_rows_[j]) s ^= message[c];
            parity[j] = s;
        }
        for (size_t j = 1; j < m_; ++j) parity[j] ^= parity[j - 1];
    } else {
        std::vector<uint64_t> msg(g0_words_, 0);
        for (size_t c = 0; c < k_; ++c) if (message[c]) msg[c / 64] |= (uint64_t)1 << (c % 64);
        for (size_t j = 0; j < m_; ++j) {
            uint64_t acc = 0;
            for (size_t w = 0; w < g0_words_; ++w) acc ^= g0_[j * g0_words_ + w] & msg[w];
            parity[j] = (uint8_t)(__builtin_popcountll(acc) & 1);
        }
    }
}

// =====================================================================================
// BER loop, src/simulation/ber.rs
// =====================================================================================
double noise_sigma(double rate, double bits_per_symbol, float ebn0_db) {   // ber.rs:300-302
    double ebn0 = std::pow(10.0, 0.1 * (double)ebn0_db);
    double esn0 = rate * bits_per_symbol * ebn0;
    return std::sqrt(0.5 / esn0);
}

bool ber_run(const SparseMatrix& h, const std::string& implementation, const std::string& puncturing,
             float ebn0_db, size_t max_iterations, uint64_t frames, uint64_t max_frame_errors,
             int nthreads, uint64_t seed, bool linear_search_send, BerCounters* out, std::string* err) {
    std::unique_ptr<Puncturer> punct;
    if (!puncturing.empty()) {
        std::vector<bool> pat;
        if (!parse_puncturing_pattern(puncturing, &pat)) { if (err) *err = "invalid puncturing pattern"; return false; }
        punct = std::make_unique<Puncturer>(pat);
    }
    std::string eerr;
    auto enc0 = Encoder::from_h(h, &eerr);
    if (!enc0) { if (err) *err = eerr; return false; }
    if (!build_decoder(implementation, h)) { if (err) *err = "invalid decoder implementation"; return false; }
    size_t n_cw = h.num_cols(), k = n_cw - h.num_rows();                      // ber.rs:247-259
    double prate = punct ? punct->rate() : 1.0;
    size_t n = (size_t)std::llround((double)n_cw / prate);
    double rate = (double)k / (double)n;
    double sigma = noise_sigma(rate, 1.0, ebn0_db);
    if (nthreads < 1) nthreads = 1;

    std::atomic<uint64_t> next_frame{0};
    std::atomic<uint64_t> frame_errors_seen{0};
    std::atomic<bool> failed{false};
    std::mutex mu;
    BerCounters total;
    auto t0 = std::chrono::steady_clock::now();
    auto worker = [&](int tid) {
        auto dec = build_decoder(implementation, h);
        dec->set_linear_search_send(linear_search_send);
        std::mt19937_64 rng(seed * 0x9E3779B97F4A7C15ULL + (uint64_t)tid * 0xD1B54A32D192ED03ULL + 1);
        std::normal_distribution<double> gauss(0.0, sigma);
        std::vector<uint8_t> msg(k), cw(n_cw), tx;
        std::vector<double> llr_tx, llr_dec;
        BerCounters c;
        while (!failed.load(std::memory_order_relaxed)) {
            if (frames > 0) {
                if (next_frame.fetch_add(1) >= frames) break;
            } else if (frame_errors_seen.load() >= max_frame_errors) break;
            // Worker::simulate, ber.rs:436-481
            uint64_t bits = 0; int nb = 0;
            for (size_t i = 0; i < k; ++i) {
                if (nb == 0) { bits = rng(); nb = 64; }
                msg[i] = (uint8_t)(bits & 1); bits >>= 1; --nb;
            }
            enc0->encode(msg.data(), cw.data());
            if (punct) { if (!punct->puncture(cw, &tx)) { failed = true; break; } } else tx = cw;
            llr_tx.resize(tx.size());
            double scale = -2.0 / (sigma * sigma);                     // modulation.rs:123-141
            for (size_t i = 0; i < tx.size(); ++i) {
                double sym = tx[i] ? 1.0 : -1.0;                        // modulation.rs:87-95
                double y = sym + gauss(rng);                            // channel.rs:60-72
                llr_tx[i] = scale * y;
            }
            const double* lp = llr_tx.data(); size_t ln = llr_tx.size();
            if (punct) { if (!punct->depuncture(llr_tx.data(), llr_tx.size(), &llr_dec)) { failed = true; break; } lp = llr_dec.data(); ln = llr_dec.size(); }
            DecodeResult r = dec->decode(lp, ln, max_iterations);
            if (r.error) { failed = true; break; }
            uint64_t be = 0;
            for (size_t i = 0; i < k; ++i) be += (msg[i] != r.codeword[i]);
            bool fe = be > 0;
            c.num_frames += 1;
            c.bit_errors += be;
            c.frame_errors += fe;
            c.false_decodes += (fe && r.success);
            c.total_iterations += r.iterations;
            if (!fe) c.correct_iterations += r.iterations;
            if (fe) frame_errors_seen.fetch_add(1);
        }
        std::lock_guard<std::mutex> g(mu);
        total.num_frames += c.num_frames; total.bit_errors += c.bit_errors;
        total.frame_errors += c.frame_errors; total.false_decodes += c.false_decodes;
        total.total_iterations += c.total_iterations; total.correct_iterations += c.correct_iterations;
    };
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
    for (auto& t : th) t.join();
    total.elapsed_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (failed) { if (err) *err = "worker failed"; return false; }
    *out = total;
    return true;
}

}  // namespace oracle
