#pragma once
// ldpc_toolbox_b200/csrc/layered_smem.cu — K3q: horizontal-layered decoding with ONE FRAME PER CTA
// and the posteriors resident in shared memory.
//
// Replaces horizontal_layered::Decoder<A>::decode (reference src/decoder/horizontal_layered.rs:49-110)
// and the update_check_messages_and_vars methods of src/decoder/arithmetic.rs (:260-292, :393-426,
// :535-574, :759-801, :1013-1066, :1197-1257) for codes whose posteriors fit in shared memory and
// whose level schedule is wide (5G-NR, CCSDS AR4JA: hundreds of independent rows per level).
//
//  * Qv (n values of f32 / f64 / i16) lives in shared memory for the whole decode: the only HBM
//    traffic of an iteration is Rcv, read once and written once = the algorithmic 2*E*s bytes.
//  * Threads of the CTA take the rows of one level (rows of a level share no column, so they commute
//    exactly with the reference's row order 0..m-1); a CTA barrier separates levels.
//  * Rcv and the column indices are stored per level in ELL order [slot j][row in level], so lane i
//    and lane i+1 touch consecutive addresses; in a quasi-cyclic code the Qv addresses of
//    consecutive rows are consecutive too (no bank conflicts).
//  * The caller's LLRs [frame][llr] are read directly (depuncture, `as f32` / quantise and the
//    raw-sign pre-check fused in) and the hard word is written directly: no ingest / emit kernels.
//  * A frame that converges ends its CTA, so a batch costs the SUM of its iterations, not
//    frames x the slowest frame of a tile.
//  * Rows of degree <= 10 run fully unrolled, register-resident rule code (rules.cuh, DT > 0).
#include <string>
#include <type_traits>

#include "decoder_impl.hpp"
#include "device_common.cuh"
#include "rules.cuh"

namespace ldpc {
namespace {

template <class Q, class R>
struct SmemLayeredParams {
    LayeredSmemGraph g;
    const void* llrs;        // [nframes][llrs_len] f32 or f64
    int in_f64;
    size_t llrs_len;
    const int* src_map;      // n entries or null
    R* rcv;                  // [nframes][ell_size]
    uint8_t* out;            // [nframes][out_stride]
    size_t out_len, out_stride;
    int32_t* iters;
    int max_iter;
    double* post;            // test hook (may be null): [nframes][n] Qv when the frame ended (horizontal_layered.rs:65-88)
};

template <class T> __device__ __forceinline__ T ld_cs(const T* p) { return __ldcs(p); }
template <> __device__ __forceinline__ int8_t ld_cs(const int8_t* p) { return (int8_t)__ldcs(reinterpret_cast<const signed char*>(p)); }
template <class T> __device__ __forceinline__ void st_cs(T* p, T v) { __stcs(p, v); }
template <> __device__ __forceinline__ void st_cs(int8_t* p, int8_t v) { __stcs(reinterpret_cast<signed char*>(p), (signed char)v); }

// one row: gather Qv (shared) and Rcv (HBM, streaming), apply the rule, scatter back
template <class F, int RULE, bool IS_I8, bool HLIM, int DT, class Q, class R>
__device__ __forceinline__ void process_row(Q* __restrict__ qs, R* __restrict__ rcv, const int* __restrict__ ell_col, int stride,
                                            int d_rt, bool first_iter, const I8Tables& tb) {
    constexpr int CAP = DT > 0 ? DT : kRuleMaxD;
    const int d = DT > 0 ? DT : d_rt;
    int col[CAP];
    Q q[CAP];
    R r[CAP];
#pragma unroll
    for (int j = 0; j < d; ++j) col[j] = __ldg(ell_col + (size_t)j * stride);
#pragma unroll
    for (int j = 0; j < d; ++j) r[j] = first_iter ? R(0) : ld_cs(rcv + (size_t)j * stride);   // horizontal_layered.rs:97-102: Rcv = 0
#pragma unroll
    for (int j = 0; j < d; ++j) q[j] = qs[col[j]];
    if (IS_I8) {
        int x[CAP], out[CAP];
#pragma unroll
        for (int j = 0; j < d; ++j) x[j] = i8_clip((int)q[j] - (int)r[j]);                    // arithmetic.rs:775, :1204
        check_rule_i8<RULE, HLIM, DT>(x, d, out, tb);
#pragma unroll
        for (int j = 0; j < d; ++j) {
            q[j] = (Q)((int)q[j] - (int)r[j] + out[j]);                                       // :797-800, :1243-1256
            r[j] = (R)out[j];
        }
    } else {
        F x[CAP], out[CAP], scratch[CAP];
#pragma unroll
        for (int j = 0; j < d; ++j) x[j] = (F)q[j] - (F)r[j];
        check_rule_float<F, RULE, DT>(x, d, out, scratch);
#pragma unroll
        for (int j = 0; j < d; ++j) {
            if (RULE == kPhi || rule_is_aminstar(RULE)) q[j] = (Q)(x[j] + out[j]);                 // :290, :1064
            else q[j] = (Q)((F)q[j] + (out[j] - (F)r[j]));                                    // :423, :571
            r[j] = (R)out[j];
        }
    }
#pragma unroll
    for (int j = 0; j < d; ++j) {
        qs[col[j]] = q[j];
        st_cs(rcv + (size_t)j * stride, r[j]);
    }
}

// Two CTAs (frames) per SM for the 4-byte and int8 arithmetics: the second frame fills the level barriers and
// the load latency of the first (ncu at 512 threads x 1 CTA: 1.8 warps per issue stalled on the barrier, 2.8
// on loads).  f64 posteriors leave room for one CTA only.
template <class F, int RULE, bool IS_I8, bool HLIM>
__global__ void __launch_bounds__(kSmemLayeredMaxThreads, sizeof(F) == 8 ? 1 : 2) layered_smem_kernel(SmemLayeredParams<typename std::conditional<IS_I8, int16_t, F>::type,
                                                      typename std::conditional<IS_I8, int8_t, F>::type> p) {
    using Q = typename std::conditional<IS_I8, int16_t, F>::type;
    using R = typename std::conditional<IS_I8, int8_t, F>::type;
    extern __shared__ __align__(16) uint8_t dyn[];
    __shared__ I8Tables tb;
    const LayeredSmemGraph& g = p.g;
    Q* qs = reinterpret_cast<Q*>(dyn);                                         // [n]
    uint32_t* raw = reinterpret_cast<uint32_t*>(dyn + (((size_t)g.n * sizeof(Q) + 15) & ~(size_t)15));   // [(n+31)/32] raw-sign bits
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31;
    const size_t frame = blockIdx.x;
    R* rcv = p.rcv + frame * g.ell_size;
    if (IS_I8) i8_tables_init(tb);

    // ---- load: depuncture (puncturing.rs:85-101), raw-sign decision (x <= 0.0), input_llr_quantize
    for (int v0 = (tid & ~31); v0 < g.n; v0 += T) {
        const int v = v0 + lane;
        double xd = 1.0;
        float xf = 1.0f;
        if (v < g.n) {
            const int src = p.src_map ? __ldg(p.src_map + v) : v;
            if (p.in_f64) xd = src >= 0 ? static_cast<const double*>(p.llrs)[frame * p.llrs_len + (size_t)src] : 0.0;
            else xf = src >= 0 ? static_cast<const float*>(p.llrs)[frame * p.llrs_len + (size_t)src] : 0.0f;
        }
        const bool neg = p.in_f64 ? xd <= 0.0 : xf <= 0.0f;
        const uint32_t word = __ballot_sync(0xffffffffu, neg && v < g.n);
        if (lane == 0) raw[v0 >> 5] = word;
        if (v < g.n) {
            if (IS_I8) qs[v] = (Q)(p.in_f64 ? quantize_i8(xd) : quantize_i8(xf));
            else qs[v] = p.in_f64 ? (Q)xd : (Q)xf;
        }
    }
    __syncthreads();

    // syndrome of the hard decisions given by hard_of(column): any unsatisfied row?
    auto unsatisfied = [&](auto hard_of) {
        uint32_t bad = 0;
        for (int l = 0; l < g.num_levels; ++l) {
            const int r0 = __ldg(g.level_ptr + l), r1 = __ldg(g.level_ptr + l + 1), stride = r1 - r0;
            const int off = __ldg(g.level_ell + l), ldeg = __ldg(g.level_deg + l);
            for (int r = r0 + tid; r < r1; r += T) {
                const int d = ldeg >= 0 ? ldeg : __ldg(g.row_deg + r);
                const int* cc = g.ell_col + off + (r - r0);
                uint32_t par = 0;
                for (int j = 0; j < d; ++j) par ^= hard_of(__ldg(cc + (size_t)j * stride));
                bad |= par;
            }
        }
        return __syncthreads_or((int)bad) != 0;
    };
    auto hard_raw = [&](int v) { return (raw[v >> 5] >> (v & 31)) & 1u; };
    auto hard_q = [&](int v) { return (uint32_t)(qs[v] <= Q(0)); };               // clip() keeps the sign (arithmetic.rs:701-703)

    int result;            // iterations, or -1
    bool from_raw = false;
    if (!unsatisfied(hard_raw)) {                                                  // horizontal_layered.rs:55-62
        result = 0;
        from_raw = true;
    } else {
        result = -1;
        for (int it = 1; it <= p.max_iter; ++it) {
            const bool first = it == 1;
            for (int l = 0; l < g.num_levels; ++l) {                               // horizontal_layered.rs:105-110
                // per-level scalars (same for every thread, L1 hits): a row's loads then depend on nothing else
                const int r0 = __ldg(g.level_ptr + l), r1 = __ldg(g.level_ptr + l + 1), stride = r1 - r0;
                const int off = __ldg(g.level_ell + l), ldeg = __ldg(g.level_deg + l);
                for (int r = r0 + tid; r < r1; r += T) {
                    const int d = ldeg >= 0 ? ldeg : __ldg(g.row_deg + r);       // quasi-cyclic codes: one degree per level
                    R* rr = rcv + off + (r - r0);
                    const int* cc = g.ell_col + off + (r - r0);
                    // O(d) rules unroll up to degree 20 (5G-NR base graph 1: rows of degree 19), the O(d^2) ones to 10
                    constexpr int kUnrollMax = (rule_is_minstar(RULE) || RULE == kTanh || sizeof(F) == 8) ? 10 : 20;
#define LDPC_ROW_CASE(D_) case D_: process_row<F, RULE, IS_I8, HLIM, (D_ <= kUnrollMax ? D_ : 0), Q, R>(qs, rr, cc, stride, d, first, tb); break;
                    switch (d) {
                        case 0: break;
                        LDPC_ROW_CASE(1) LDPC_ROW_CASE(2) LDPC_ROW_CASE(3) LDPC_ROW_CASE(4) LDPC_ROW_CASE(5)
                        LDPC_ROW_CASE(6) LDPC_ROW_CASE(7) LDPC_ROW_CASE(8) LDPC_ROW_CASE(9) LDPC_ROW_CASE(10)
                        LDPC_ROW_CASE(11) LDPC_ROW_CASE(12) LDPC_ROW_CASE(13) LDPC_ROW_CASE(14) LDPC_ROW_CASE(15)
                        LDPC_ROW_CASE(16) LDPC_ROW_CASE(17) LDPC_ROW_CASE(18) LDPC_ROW_CASE(19) LDPC_ROW_CASE(20)
                        default: process_row<F, RULE, IS_I8, HLIM, 0, Q, R>(qs, rr, cc, stride, d, first, tb); break;
                    }
#undef LDPC_ROW_CASE
                }
                __syncthreads();
            }
            if (!unsatisfied(hard_q)) { result = it; break; }
        }
    }
    if (tid == 0) p.iters[frame] = result;
    if (p.post)
        for (int v = tid; v < g.n; v += T) p.post[frame * (size_t)g.n + v] = (double)qs[v];
    for (size_t v = tid; v < p.out_len; v += T)
        p.out[frame * p.out_stride + v] = (uint8_t)(from_raw ? hard_raw((int)v) : hard_q((int)v));
}

template <class F, int RULE, bool IS_I8, bool HLIM>
bool launch_t(const LayeredSmemLaunch& L, cudaStream_t stream) {
    using Q = typename std::conditional<IS_I8, int16_t, F>::type;
    using R = typename std::conditional<IS_I8, int8_t, F>::type;
    SmemLayeredParams<Q, R> p;
    p.g = L.graph; p.llrs = L.llrs; p.in_f64 = L.in_f64 ? 1 : 0; p.llrs_len = L.llrs_len; p.src_map = L.src_map;
    p.rcv = static_cast<R*>(L.rcv); p.out = L.out; p.out_len = L.out_len; p.out_stride = L.out_stride; p.iters = L.iters;
    p.max_iter = L.max_iter;
    p.post = L.post;
    const size_t smem = layered_smem_bytes(L.graph.n, L.is_f64, L.is_i8);
    // per device and cheap: set on every launch (one process may drive several GPUs)
    LDPC_CUDA_CHECK(cudaFuncSetAttribute(layered_smem_kernel<F, RULE, IS_I8, HLIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    layered_smem_kernel<F, RULE, IS_I8, HLIM><<<dim3((unsigned)L.nframes), dim3((unsigned)L.threads), smem, stream>>>(p);
    LDPC_CUDA_CHECK(cudaGetLastError());
    return true;
}

}  // namespace
}  // namespace ldpc
