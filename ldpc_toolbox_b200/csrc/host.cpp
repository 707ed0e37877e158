// ldpc_toolbox_b200/csrc/host.cpp — see host.hpp for the reference interfaces mirrored here.
#include "host.hpp"

#include <algorithm>
#include <cctype>
#include <cstring>

namespace ldpc {

// ------------------------------------------------------------------------------------------------
// alist -> edge-index layout.  Behaviour follows reference src/sparse.rs:352-389: only line 1 and
// the ncols column lines (lines 5..) are read, zeros are padding, duplicates are ignored
// (SparseMatrix::insert, sparse.rs:114-119).
// ------------------------------------------------------------------------------------------------
namespace {
struct LineCursor {
    const char* p;
    const char* end;
    bool done = false;      // true once the text after the last '\n' has been handed out
    // next '\n'-delimited line (the text after the final '\n' counts as one more, possibly empty, line)
    bool next(const char** b, const char** e) {
        if (done) return false;
        const char* q = static_cast<const char*>(memchr(p, '\n', size_t(end - p)));
        *b = p;
        if (q) { *e = q; p = q + 1; }
        else { *e = end; done = true; }
        return true;
    }
};

// whitespace-separated unsigned tokens of one line; returns false on a malformed token
template <class F>
bool for_each_number(const char* b, const char* e, F&& f) {
    while (b < e) {
        while (b < e && isspace((unsigned char)*b)) ++b;
        if (b == e) break;
        const char* t = b;
        while (b < e && !isspace((unsigned char)*b)) ++b;
        if (*t == '+') ++t;
        if (t == b) return false;
        unsigned long long v = 0;
        for (const char* c = t; c < b; ++c) {
            if (*c < '0' || *c > '9') return false;
            if (v > (~0ULL - 9) / 10) return false;
            v = v * 10 + (unsigned)(*c - '0');
        }
        if (!f(v)) return false;
    }
    return true;
}
}  // namespace

bool Graph::from_alist(const std::string& text, Graph* out, std::string* err) {
    auto fail = [&](const char* m) { if (err) *err = m; return false; };
    LineCursor cur{text.data(), text.data() + text.size()};
    const char *b, *e;
    if (!cur.next(&b, &e)) return fail("alist first line not found");
    unsigned long long dims[2];
    int nd = 0;
    bool bad_number = false;
    {
        // only the first two tokens are parsed by the reference
        const char* q = b;
        while (nd < 2) {
            while (q < e && isspace((unsigned char)*q)) ++q;
            if (q == e) break;
            const char* t = q;
            while (q < e && !isspace((unsigned char)*q)) ++q;
            unsigned long long v = 0;
            bool ok = for_each_number(t, q, [&](unsigned long long x) { v = x; return true; });
            if (!ok) { bad_number = true; break; }
            dims[nd++] = v;
        }
    }
    if (bad_number) return fail(nd == 0 ? "ncols is not a number" : "nrows is not a number");
    if (nd < 2) return fail("alist first line does not contain enough elements");
    if (dims[0] > 0x3fffffffULL || dims[1] > 0x3fffffffULL) return fail("matrix too large");
    const int ncols = (int)dims[0], nrows = (int)dims[1];
    for (int i = 0; i < 3; ++i) cur.next(&b, &e);   // max weights, column weights, row weights: ignored

    // every column needs a line of its own, so a header that promises more columns than the text has bytes is
    // refused before anything is sized by it
    if ((size_t)ncols > text.size()) return fail("alist does not contain expected number of lines");
    std::vector<int> col_ptr((size_t)ncols + 1, 0), col_row;
    for (int c = 0; c < ncols; ++c) {
        if (!cur.next(&b, &e)) return fail("alist does not contain expected number of lines");
        const size_t start = col_row.size();
        bool range_ok = true;
        bool ok = for_each_number(b, e, [&](unsigned long long v) {
            if (v == 0) return true;                       // padding
            if (v > (unsigned long long)nrows) { range_ok = false; return false; }
            int r = (int)(v - 1);
            for (size_t i = start; i < col_row.size(); ++i)
                if (col_row[i] == r) return true;          // duplicate entry: insert() is a no-op
            col_row.push_back(r);
            return true;
        });
        if (!range_ok) return fail("row index out of range");
        if (!ok) return fail("row value is not a number");
        if (col_row.size() > 0x7fffffffULL) return fail("matrix too large");
        col_ptr[(size_t)c + 1] = (int)col_row.size();
    }

    // rows by counting: columns are visited in increasing order, so every row's entries come out sorted by
    // column, and the slot an entry lands in is its edge index
    Graph g;
    g.n = ncols; g.m = nrows; g.E = (int)col_row.size();
    g.row_ptr.assign((size_t)nrows + 1, 0);
    for (int r : col_row) ++g.row_ptr[(size_t)r + 1];
    g.max_row_deg = 0; g.min_row_deg = nrows ? 1 << 30 : 0;
    for (int r = 0; r < nrows; ++r) {
        const int d = g.row_ptr[(size_t)r + 1];
        g.max_row_deg = std::max(g.max_row_deg, d);
        g.min_row_deg = std::min(g.min_row_deg, d);
        g.row_ptr[(size_t)r + 1] += g.row_ptr[(size_t)r];
    }
    g.col_idx.resize((size_t)g.E);
    g.col_edge.resize((size_t)g.E);
    std::vector<int> fill(g.row_ptr.begin(), g.row_ptr.end() - 1);
    for (int c = 0; c < ncols; ++c) {
        for (int p = col_ptr[(size_t)c]; p < col_ptr[(size_t)c + 1]; ++p) {
            const int slot = fill[(size_t)col_row[(size_t)p]]++;
            g.col_idx[(size_t)slot] = c;
            g.col_edge[(size_t)p] = slot;
        }
        g.max_col_deg = std::max(g.max_col_deg, col_ptr[(size_t)c + 1] - col_ptr[(size_t)c]);
    }
    g.col_ptr = std::move(col_ptr);
    g.col_row = std::move(col_row);
    *out = std::move(g);
    return true;
}

// ------------------------------------------------------------------------------------------------
// Implementation names, reference src/decoder/factory.rs:240-277
// ------------------------------------------------------------------------------------------------
namespace {
std::vector<DecoderImplementation> make_table() {
    std::vector<DecoderImplementation> t;
    auto add = [&](const std::string& name, Rule r, Dtype d, Schedule s, bool j = false, bool h = false, bool c = false) {
        DecoderImplementation x;
        x.name = name; x.rule = r; x.dtype = d; x.schedule = s; x.jones = j; x.hardlimit = h; x.deg1clip = c;
        t.push_back(x);
    };
    const Schedule F = Schedule::Flooding, H = Schedule::HorizontalLayered;
    struct FR { const char* n; Rule r; };
    const FR float_rules[] = {{"Phi", Rule::Phi}, {"Tanh", Rule::Tanh}, {"Minstarapprox", Rule::Minstarapprox}};
    for (const auto& fr : float_rules) {
        add(std::string(fr.n) + "f64", fr.r, Dtype::F64, F);
        add(std::string(fr.n) + "f32", fr.r, Dtype::F32, F);
    }
    auto add_i8_family = [&](const char* prefix, Rule r) {
        for (int c = 0; c < 2; ++c)
            for (int h = 0; h < 2; ++h)
                for (int j = 0; j < 2; ++j) {
                    std::string nm = prefix;
                    if (j) nm += "Jones";
                    if (h) nm += "PartialHardLimit";
                    if (c) nm += "Deg1Clip";
                    add(nm, r, Dtype::I8, F, j, h, c);
                }
    };
    add_i8_family("Minstarapproxi8", Rule::Minstarapprox);
    add("Aminstarf64", Rule::Aminstar, Dtype::F64, F);
    add("Aminstarf32", Rule::Aminstar, Dtype::F32, F);
    add_i8_family("Aminstari8", Rule::Aminstar);
    for (const auto& fr : float_rules) {
        add(std::string("HL") + fr.n + "f64", fr.r, Dtype::F64, H);
        add(std::string("HL") + fr.n + "f32", fr.r, Dtype::F32, H);
    }
    add("HLMinstarapproxi8", Rule::Minstarapprox, Dtype::I8, H);
    add("HLMinstarapproxi8PartialHardLimit", Rule::Minstarapprox, Dtype::I8, H, false, true);
    add("HLAminstarf64", Rule::Aminstar, Dtype::F64, H);
    add("HLAminstarf32", Rule::Aminstar, Dtype::F32, H);
    add("HLAminstari8", Rule::Aminstar, Dtype::I8, H);
    add("HLAminstari8PartialHardLimit", Rule::Aminstar, Dtype::I8, H, false, true);
    return t;
}
const std::vector<DecoderImplementation>& table() {
    static const std::vector<DecoderImplementation> t = make_table();
    return t;
}
}  // namespace

bool DecoderImplementation::parse(const std::string& s, DecoderImplementation* out) {
    for (const auto& x : table())
        if (x.name == s) { *out = x; return true; }
    return false;
}

const std::vector<std::string>& DecoderImplementation::all_names() {
    static const std::vector<std::string> names = [] {
        std::vector<std::string> v;
        for (const auto& x : table()) v.push_back(x.name);
        return v;
    }();
    return names;
}

// ------------------------------------------------------------------------------------------------
// Puncturing (reference src/cli/ber.rs:219-229, src/simulation/puncturing.rs:47-110)
// ------------------------------------------------------------------------------------------------
bool parse_puncturing_pattern(const std::string& s, std::vector<bool>* out) {
    out->clear();
    size_t b = 0;
    for (;;) {
        size_t c = s.find(',', b);
        size_t len = (c == std::string::npos ? s.size() : c) - b;
        if (len != 1 || (s[b] != '0' && s[b] != '1')) return false;
        out->push_back(s[b] == '1');
        if (c == std::string::npos) return true;
        b = c + 1;
    }
}

Puncturer::Puncturer(const std::vector<bool>& p) : pattern(p) {
    num_trues = (size_t)std::count(p.begin(), p.end(), true);
}

bool Puncturer::depuncture_map(size_t punctured_len, size_t n_cw, std::vector<int>* map) const {
    if (num_trues == 0 || punctured_len % num_trues != 0) return false;
    const size_t bs = punctured_len / num_trues;
    if (bs * pattern.size() != n_cw) return false;
    map->assign(n_cw, -1);
    size_t j = 0;
    for (size_t blk = 0; blk < pattern.size(); ++blk) {
        if (!pattern[blk]) continue;
        for (size_t i = 0; i < bs; ++i) (*map)[blk * bs + i] = (int)(j * bs + i);
        ++j;
    }
    return true;
}

bool Puncturer::puncture_map(size_t n_cw, std::vector<int>* kept) const {
    if (n_cw % pattern.size() != 0) return false;
    const size_t bs = n_cw / pattern.size();
    kept->clear();
    for (size_t blk = 0; blk < pattern.size(); ++blk)
        if (pattern[blk])
            for (size_t i = 0; i < bs; ++i) kept->push_back((int)(blk * bs + i));
    return true;
}

// ------------------------------------------------------------------------------------------------
// Encoder plan (reference src/encoder.rs:59-120, src/encoder/staircase.rs:3-24)
// ------------------------------------------------------------------------------------------------
bool EncoderPlan::from_graph(const Graph& g, EncoderPlan* out, std::string* err) {
    if (g.m > g.n) { if (err) *err = "more rows than columns"; return false; }
    EncoderPlan p;
    p.n = g.n; p.m = g.m; p.k = g.n - g.m;
    // staircase test: every one in the last m columns sits on the diagonal or the sub-diagonal and
    // there are exactly 2m-1 of them
    long parity_ones = 0;
    bool stair = g.m > 0;
    for (int r = 0; r < g.m && stair; ++r)
        for (int q = g.row_ptr[(size_t)r]; q < g.row_ptr[(size_t)r + 1]; ++q) {
            int c = g.col_idx[(size_t)q];
            if (c < p.k) continue;
            int d = c - p.k;
            if (d != r && d != r - 1) { stair = false; break; }
            ++parity_ones;
        }
    stair = stair && parity_ones == 2L * g.m - 1;
    p.staircase = stair;
    if (stair) {
        p.h0_ptr.assign((size_t)g.m + 1, 0);
        for (int r = 0; r < g.m; ++r) {
            for (int q = g.row_ptr[(size_t)r]; q < g.row_ptr[(size_t)r + 1]; ++q)
                if (g.col_idx[(size_t)q] < p.k) p.h0_idx.push_back(g.col_idx[(size_t)q]);
            p.h0_ptr[(size_t)r + 1] = (int)p.h0_idx.size();
        }
        *out = std::move(p);
        return true;
    }
    // dense path: eliminate on [H1 | H0], all rows packed 64 columns per word
    const int W = (g.n + 63) / 64;
    std::vector<uint64_t> a((size_t)g.m * (size_t)W, 0);
    auto at = [&](int r) { return a.data() + (size_t)r * (size_t)W; };
    for (int r = 0; r < g.m; ++r)
        for (int q = g.row_ptr[(size_t)r]; q < g.row_ptr[(size_t)r + 1]; ++q) {
            int c = g.col_idx[(size_t)q];
            int t = c < p.k ? c + g.m : c - p.k;
            at(r)[t >> 6] |= 1ULL << (t & 63);
        }
    for (int j = 0; j < g.m; ++j) {
        int piv = -1;
        for (int r = j; r < g.m; ++r)
            if ((at(r)[j >> 6] >> (j & 63)) & 1) { piv = r; break; }
        if (piv < 0) {
            if (err) *err = "the square matrix formed by the last columns of the parity check is not invertible";
            return false;
        }
        if (piv != j) std::swap_ranges(at(j), at(j) + W, at(piv));
        const uint64_t* pj = at(j);
        for (int r = 0; r < g.m; ++r) {
            if (r == j || !((at(r)[j >> 6] >> (j & 63)) & 1)) continue;
            uint64_t* pr = at(r);
            for (int w = j >> 6; w < W; ++w) pr[w] ^= pj[w];
        }
    }
    p.words = (p.k + 63) / 64;
    p.g0.assign((size_t)g.m * (size_t)p.words, 0);
    for (int r = 0; r < g.m; ++r)
        for (int c = 0; c < p.k; ++c) {
            int t = g.m + c;
            if ((at(r)[t >> 6] >> (t & 63)) & 1) p.g0[(size_t)r * (size_t)p.words + (size_t)(c >> 6)] |= 1ULL << (c & 63);
        }
    *out = std::move(p);
    return true;
}

void EncoderPlan::encode(const uint8_t* msg01, uint8_t* cw01) const {
    memcpy(cw01, msg01, (size_t)k);
    uint8_t* par = cw01 + k;
    if (staircase) {
        uint8_t run = 0;
        for (int r = 0; r < m; ++r) {
            uint8_t s = 0;
            for (int q = h0_ptr[(size_t)r]; q < h0_ptr[(size_t)r + 1]; ++q) s ^= msg01[h0_idx[(size_t)q]];
            run ^= s;                 // running XOR = accumulate (encoder.rs:112-116)
            par[r] = run;
        }
        return;
    }
    std::vector<uint64_t> mw((size_t)words, 0);
    for (int c = 0; c < k; ++c)
        if (msg01[c]) mw[(size_t)(c >> 6)] |= 1ULL << (c & 63);
    for (int r = 0; r < m; ++r) {
        uint64_t acc = 0;
        const uint64_t* row = g0.data() + (size_t)r * (size_t)words;
        for (int w = 0; w < words; ++w) acc ^= row[w] & mw[(size_t)w];
        par[r] = (uint8_t)(__builtin_popcountll(acc) & 1);
    }
}

}  // namespace ldpc
