// ldpc_toolbox_b200/csrc/generic_bp.cu — K2 (flooding, float rules) and K3 (horizontal layered, all
// rules): correctness-first kernels for the 20 implementations that are not the packed int8
// flooding path (flood_i8.cu).
//
//   K2 replaces flooding::Decoder<A>::decode for A in {Phi, Tanh, Minstarapprox, Aminstar} x {f64, f32}
//      reference src/decoder/flooding.rs:51-125, src/decoder/arithmetic.rs:140-156 (variable node)
//   K3 replaces horizontal_layered::Decoder<A>::decode for the 12 HL* implementations
//      reference src/decoder/horizontal_layered.rs:49-110 and the update_check_messages_and_vars
//      methods of src/decoder/arithmetic.rs (:260-292, :393-426, :535-574, :759-801, :1013-1066,
//      :1197-1257)
//
// Layout: 128-frame tiles, frame index fastest ([node][128] values); a lane owns 4 consecutive
// frames and runs the reference's per-frame arithmetic on them one after the other, in the
// reference's order.  One CTA owns a tile for the whole decode.
//
// The layered schedule is sequential over rows inside a frame.  Rows whose column supports are
// disjoint commute exactly, so the host builds a level schedule (a row's level is one more than the
// highest level of any earlier row sharing a column with it); rows of one level run on different
// warps, levels are separated by a CTA barrier, and the result is identical to the reference's
// row order 0..m-1.  (5G-NR: 384 rows per level; DVB-S2: the staircase chains every row to the
// next, so its layered decoders run one row at a time.)
#include <string>
#include <type_traits>

#include "decoder_impl.hpp"
#include "device_common.cuh"
#include "rules.cuh"

namespace ldpc {
namespace {

constexpr int kGWarps = 8;

// ---- 4-frame vector access ([node][128] arrays, lane owns frames 4*lane .. 4*lane+3) -----------
template <class T> struct V4 { T v[4]; };

template <class T> __device__ __forceinline__ V4<T> ld4(const T* base, size_t node, int lane);
template <> __device__ __forceinline__ V4<float> ld4(const float* base, size_t node, int lane) {
    float4 t = *reinterpret_cast<const float4*>(base + node * kTileFrames + lane * 4);
    return {{t.x, t.y, t.z, t.w}};
}
template <> __device__ __forceinline__ V4<double> ld4(const double* base, size_t node, int lane) {
    const double2* p = reinterpret_cast<const double2*>(base + node * kTileFrames + lane * 4);
    double2 a = p[0], b = p[1];
    return {{a.x, a.y, b.x, b.y}};
}
template <> __device__ __forceinline__ V4<int8_t> ld4(const int8_t* base, size_t node, int lane) {
    uint32_t t = *reinterpret_cast<const uint32_t*>(base + node * kTileFrames + lane * 4);
    return {{(int8_t)t, (int8_t)(t >> 8), (int8_t)(t >> 16), (int8_t)(t >> 24)}};
}
template <> __device__ __forceinline__ V4<int16_t> ld4(const int16_t* base, size_t node, int lane) {
    uint2 t = *reinterpret_cast<const uint2*>(base + node * kTileFrames + lane * 4);
    return {{(int16_t)t.x, (int16_t)(t.x >> 16), (int16_t)t.y, (int16_t)(t.y >> 16)}};
}
template <class T> __device__ __forceinline__ void st4(T* base, size_t node, int lane, const V4<T>& x);
template <> __device__ __forceinline__ void st4(float* base, size_t node, int lane, const V4<float>& x) {
    *reinterpret_cast<float4*>(base + node * kTileFrames + lane * 4) = make_float4(x.v[0], x.v[1], x.v[2], x.v[3]);
}
template <> __device__ __forceinline__ void st4(double* base, size_t node, int lane, const V4<double>& x) {
    double2* p = reinterpret_cast<double2*>(base + node * kTileFrames + lane * 4);
    p[0] = make_double2(x.v[0], x.v[1]);
    p[1] = make_double2(x.v[2], x.v[3]);
}
template <> __device__ __forceinline__ void st4(int8_t* base, size_t node, int lane, const V4<int8_t>& x) {
    uint32_t t = (uint32_t)(uint8_t)x.v[0] | (uint32_t)(uint8_t)x.v[1] << 8 | (uint32_t)(uint8_t)x.v[2] << 16 | (uint32_t)(uint8_t)x.v[3] << 24;
    *reinterpret_cast<uint32_t*>(base + node * kTileFrames + lane * 4) = t;
}
template <> __device__ __forceinline__ void st4(int16_t* base, size_t node, int lane, const V4<int16_t>& x) {
    uint2 t;
    t.x = (uint32_t)(uint16_t)x.v[0] | (uint32_t)(uint16_t)x.v[1] << 16;
    t.y = (uint32_t)(uint16_t)x.v[2] | (uint32_t)(uint16_t)x.v[3] << 16;
    *reinterpret_cast<uint2*>(base + node * kTileFrames + lane * 4) = t;
}

// ---- shared early-termination bookkeeping ------------------------------------------------------
struct StopState {
    uint32_t unsat[kLanes];
    uint32_t done[kLanes];
};

// =================================================================================================
// K2: flooding, float rules
// =================================================================================================
template <class F>
struct FloodFloatParams {
    DeviceGraph g;
    F* msg;                 // [tiles][E][128]   v->c / c->v in place, row-major edge order
    uint8_t* hbit;          // [tiles][E][32]    hard decision of the edge's variable (4 bits per lane)
    const F* in;            // [tiles][n][128]   channel LLRs (`llr as F`)
    const uint8_t* raw0;    // [tiles][n][32]    raw-sign hard decisions
    uint8_t* final_hard;    // [tiles][n][32]
    int32_t* iters;         // [tiles*128]
    int max_iter;
};

// One check of a 128-frame tile.  DT > 0: compile-time degree, everything in registers.  The four
// frames of a lane are processed one after the other by rotating the components of the 4-vectors
// (component 0 is consumed, the result re-enters as component 3), so the frame loop stays rolled
// without ever indexing a register array with a run-time value.
template <class F, int RULE, int DT>
__device__ __forceinline__ void flood_check_row(F* __restrict__ msg, size_t e0, int d_rt, int lane) {
    constexpr int CAP = DT > 0 ? DT : kRuleMaxD;
    const int d = DT > 0 ? DT : d_rt;
    V4<F> xs[CAP];
#pragma unroll
    for (int j = 0; j < d; ++j) xs[j] = ld4<F>(msg, e0 + j, lane);
#pragma unroll 1
    for (int f = 0; f < 4; ++f) {
        F x[CAP], out[CAP], scratch[CAP];
#pragma unroll
        for (int j = 0; j < d; ++j) x[j] = xs[j].v[0];
        check_rule_float<F, RULE, DT>(x, d, out, scratch);
#pragma unroll
        for (int j = 0; j < d; ++j) { xs[j].v[0] = xs[j].v[1]; xs[j].v[1] = xs[j].v[2]; xs[j].v[2] = xs[j].v[3]; xs[j].v[3] = out[j]; }
    }
#pragma unroll
    for (int j = 0; j < d; ++j) st4<F>(msg, e0 + j, lane, xs[j]);
}

template <class F, int RULE>
__global__ void __launch_bounds__(kGWarps * 32) flood_float_kernel(FloodFloatParams<F> p) {
    __shared__ StopState st;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t tile = blockIdx.x;
    const DeviceGraph& g = p.g;
    F* msg = p.msg + tile * (size_t)g.E * kTileFrames;
    uint8_t* hbit = p.hbit + tile * (size_t)g.E * kLanes;
    const F* in = p.in + tile * (size_t)g.n * kTileFrames;
    const uint8_t* raw0 = p.raw0 + tile * (size_t)g.n * kLanes;
    uint8_t* fin = p.final_hard + tile * (size_t)g.n * kLanes;
    int32_t* iters = p.iters + tile * kTileFrames;
    if (threadIdx.x < kLanes) { st.unsat[threadIdx.x] = 0; st.done[threadIdx.x] = 0; }

    // flooding.rs:88-100
    for (int v = warp; v < g.n; v += kGWarps) {
        V4<F> w = ld4<F>(in, (size_t)v, lane);
        uint8_t hb = raw0[(size_t)v * kLanes + lane];
        for (int q = __ldg(g.col_ptr + v); q < __ldg(g.col_ptr + v + 1); ++q) {
            size_t e = (size_t)__ldg(g.col_edge + q);
            st4<F>(msg, e, lane, w);
            hbit[e * kLanes + lane] = hb;
        }
    }
    __syncthreads();

    for (int it = 1;; ++it) {
        const bool last = it > p.max_iter;
        uint32_t synd = 0;
        for (int c = warp; c < g.m; c += kGWarps) {                      // flooding.rs:102-109
            const int e0 = __ldg(g.row_ptr + c), d = __ldg(g.row_ptr + c + 1) - e0;
            uint32_t hb = 0;
            for (int j = 0; j < d; ++j) hb ^= hbit[(size_t)(e0 + j) * kLanes + lane];
            synd |= hb;
            if (last || d == 0) continue;
#define LDPC_CHK_CASE(D_) case D_: flood_check_row<F, RULE, D_>(msg, (size_t)e0, d, lane); break;
            switch (d) {
                LDPC_CHK_CASE(1) LDPC_CHK_CASE(2) LDPC_CHK_CASE(3) LDPC_CHK_CASE(4) LDPC_CHK_CASE(5) LDPC_CHK_CASE(6)
                LDPC_CHK_CASE(7) LDPC_CHK_CASE(8) LDPC_CHK_CASE(9) LDPC_CHK_CASE(10)
                default: flood_check_row<F, RULE, 0>(msg, (size_t)e0, d, lane); break;
            }
#undef LDPC_CHK_CASE
        }
        if (synd) atomicOr(&st.unsat[lane], synd);
        __syncthreads();
        const uint32_t unsat = st.unsat[lane], done = st.done[lane];
        uint32_t stop = ~unsat & ~done & 0xfu, fail = 0;
        if (last) { fail = unsat & ~done & 0xfu; stop |= fail; }
        const int any = __syncthreads_or(stop != 0);
        if (warp == 0) st.unsat[lane] = 0;
        if (any) {
            if (stop) {
                for (int v = warp; v < g.n; v += kGWarps) {
                    size_t o = (size_t)v * kLanes + lane;
                    int p0 = __ldg(g.col_ptr + v), p1 = __ldg(g.col_ptr + v + 1);
                    uint32_t hb;
                    if (p1 > p0) hb = hbit[(size_t)__ldg(g.col_edge + p0) * kLanes + lane];
                    else if (it == 1) hb = raw0[o];
                    else {
                        V4<F> w = ld4<F>(in, (size_t)v, lane);
                        hb = 0;
#pragma unroll
                        for (int b = 0; b < 4; ++b) hb |= (uint32_t)(w.v[b] <= F(0)) << b;
                    }
                    fin[o] = (uint8_t)((fin[o] & ~stop) | (hb & stop));
                }
            }
            if (warp == 0) {
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (stop >> b & 1) iters[lane * 4 + b] = (fail >> b & 1) ? -1 : it - 1;
                st.done[lane] = done | stop;
            }
        }
        const int all = __syncthreads_and(((done | stop) & 0xfu) == 0xfu);
        if (all || last) break;

        for (int v = warp; v < g.n; v += kGWarps) {                      // flooding.rs:111-125, arithmetic.rs:140-156
            const int p0 = __ldg(g.col_ptr + v), d = __ldg(g.col_ptr + v + 1) - p0;
            V4<F> inp = ld4<F>(in, (size_t)v, lane);
            F sum[4] = {F(0), F(0), F(0), F(0)};
            for (int j = 0; j < d; ++j) {
                V4<F> c = ld4<F>(msg, (size_t)__ldg(g.col_edge + p0 + j), lane);
#pragma unroll
                for (int f = 0; f < 4; ++f) sum[f] += c.v[f];
            }
            F llr[4];
            uint32_t hb = 0;
#pragma unroll
            for (int f = 0; f < 4; ++f) { llr[f] = inp.v[f] + sum[f]; hb |= (uint32_t)(llr[f] <= F(0)) << f; }
            for (int j = 0; j < d; ++j) {
                size_t e = (size_t)__ldg(g.col_edge + p0 + j);
                V4<F> c = ld4<F>(msg, e, lane);
#pragma unroll
                for (int f = 0; f < 4; ++f) c.v[f] = llr[f] - c.v[f];
                st4<F>(msg, e, lane, c);
                hbit[e * kLanes + lane] = (uint8_t)hb;
            }
        }
        __syncthreads();
    }
}

// =================================================================================================
// K3: horizontal layered
// =================================================================================================
template <class Q, class R>
struct LayeredParams {
    DeviceGraph g;
    const int* level_ptr;   // num_levels+1
    const int* level_rows;  // m, rows grouped by level, ascending inside a level
    int num_levels;
    Q* qv;                  // [tiles][n][128]   posteriors (VarLlr)
    R* rcv;                 // [tiles][E][128]   check->variable messages
    const uint8_t* raw0;    // [tiles][n][32]
    uint8_t* final_hard;    // [tiles][n][32]
    int32_t* iters;
    int max_iter;
};

// syndrome of 4-bit-per-lane hard decisions produced by `hard_of(v)`
template <class HardOf>
__device__ __forceinline__ uint32_t syndrome_pass(const DeviceGraph& g, int warp, HardOf hard_of) {
    uint32_t synd = 0;
    for (int c = warp; c < g.m; c += kGWarps) {
        const int e0 = __ldg(g.row_ptr + c), e1 = __ldg(g.row_ptr + c + 1);
        uint32_t hb = 0;
        for (int e = e0; e < e1; ++e) hb ^= hard_of(__ldg(g.col_idx + e));
        synd |= hb;
    }
    return synd;
}

// One row of a 128-frame tile (see flood_check_row for the component rotation).
template <class F, int RULE, bool IS_I8, bool HLIM, int DT, class Q, class R>
__device__ __forceinline__ void layered_tile_row(Q* __restrict__ qv, R* __restrict__ rcv, const int* __restrict__ col_idx, int e0,
                                                 int d_rt, int lane, const I8Tables& tb) {
    constexpr int CAP = DT > 0 ? DT : kRuleMaxD;
    const int d = DT > 0 ? DT : d_rt;
    int col[CAP];
    V4<Q> qs[CAP];
    V4<R> rs[CAP];
#pragma unroll
    for (int j = 0; j < d; ++j) col[j] = __ldg(col_idx + e0 + j);
#pragma unroll
    for (int j = 0; j < d; ++j) {
        qs[j] = ld4<Q>(qv, (size_t)col[j], lane);
        rs[j] = ld4<R>(rcv, (size_t)(e0 + j), lane);
    }
#pragma unroll 1
    for (int f = 0; f < 4; ++f) {
        Q qn[CAP];
        R rn[CAP];
        if (IS_I8) {
            int x[CAP], out[CAP];
#pragma unroll
            for (int j = 0; j < d; ++j) x[j] = i8_clip((int)qs[j].v[0] - (int)rs[j].v[0]);   // arithmetic.rs:775, :1204
            check_rule_i8<RULE, HLIM, DT>(x, d, out, tb);
#pragma unroll
            for (int j = 0; j < d; ++j) {
                // :797-800 and :1243-1256 are the same integer update
                qn[j] = (Q)((int)qs[j].v[0] - (int)rs[j].v[0] + out[j]);
                rn[j] = (R)out[j];
            }
        } else {
            F x[CAP], out[CAP], scratch[CAP];
#pragma unroll
            for (int j = 0; j < d; ++j) x[j] = (F)qs[j].v[0] - (F)rs[j].v[0];
            check_rule_float<F, RULE, DT>(x, d, out, scratch);
#pragma unroll
            for (int j = 0; j < d; ++j) {
                if (RULE == kPhi || RULE == kAminstar) qn[j] = (Q)(x[j] + out[j]);                      // :290, :1064
                else qn[j] = (Q)((F)qs[j].v[0] + (out[j] - (F)rs[j].v[0]));                             // :423, :571
                rn[j] = (R)out[j];
            }
        }
#pragma unroll
        for (int j = 0; j < d; ++j) {
            qs[j].v[0] = qs[j].v[1]; qs[j].v[1] = qs[j].v[2]; qs[j].v[2] = qs[j].v[3]; qs[j].v[3] = qn[j];
            rs[j].v[0] = rs[j].v[1]; rs[j].v[1] = rs[j].v[2]; rs[j].v[2] = rs[j].v[3]; rs[j].v[3] = rn[j];
        }
    }
#pragma unroll
    for (int j = 0; j < d; ++j) {
        st4<Q>(qv, (size_t)col[j], lane, qs[j]);
        st4<R>(rcv, (size_t)(e0 + j), lane, rs[j]);
    }
}

template <class F, int RULE, bool IS_I8, bool HLIM>
__global__ void __launch_bounds__(kGWarps * 32)
layered_kernel(LayeredParams<typename std::conditional<IS_I8, int16_t, F>::type, typename std::conditional<IS_I8, int8_t, F>::type> p) {
    using Q = typename std::conditional<IS_I8, int16_t, F>::type;
    using R = typename std::conditional<IS_I8, int8_t, F>::type;
    __shared__ StopState st;
    __shared__ I8Tables tb;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t tile = blockIdx.x;
    const DeviceGraph& g = p.g;
    Q* qv = p.qv + tile * (size_t)g.n * kTileFrames;
    R* rcv = p.rcv + tile * (size_t)g.E * kTileFrames;
    const uint8_t* raw0 = p.raw0 + tile * (size_t)g.n * kLanes;
    uint8_t* fin = p.final_hard + tile * (size_t)g.n * kLanes;
    int32_t* iters = p.iters + tile * kTileFrames;
    if (threadIdx.x < kLanes) { st.unsat[threadIdx.x] = 0; st.done[threadIdx.x] = 0; }
    if (IS_I8) i8_tables_init(tb);

    auto hard_q = [&](int v) {                  // llr_hard_decision(var_llr_to_llr(Qv)); clip keeps the sign
        V4<Q> q = ld4<Q>(qv, (size_t)v, lane);
        uint32_t hb = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) hb |= (uint32_t)(q.v[b] <= Q(0)) << b;
        return hb;
    };
    auto snapshot = [&](uint32_t stop, bool from_raw) {
        for (int v = warp; v < g.n; v += kGWarps) {
            size_t o = (size_t)v * kLanes + lane;
            uint32_t hb = from_raw ? (uint32_t)raw0[o] : hard_q(v);
            fin[o] = (uint8_t)((fin[o] & ~stop) | (hb & stop));
        }
    };
    // horizontal_layered.rs:97-102: Rcv = 0 (Qv was initialised by the ingest kernel, :90-96)
    {
        V4<R> z = {{R(0), R(0), R(0), R(0)}};
        for (int e = warp; e < g.E; e += kGWarps) st4<R>(rcv, (size_t)e, lane, z);
    }
    __syncthreads();

    for (int it = 0;; ++it) {
        // ---- syndrome of the current hard decisions (it == 0: raw LLR signs, :55-62)
        uint32_t synd = it == 0 ? syndrome_pass(g, warp, [&](int v) { return (uint32_t)raw0[(size_t)v * kLanes + lane]; })
                                : syndrome_pass(g, warp, hard_q);
        if (synd) atomicOr(&st.unsat[lane], synd);
        __syncthreads();
        const bool last = it >= p.max_iter;
        const uint32_t unsat = st.unsat[lane], done = st.done[lane];
        uint32_t stop = ~unsat & ~done & 0xfu, fail = 0;
        if (last) { fail = unsat & ~done & 0xfu; stop |= fail; }
        const int any = __syncthreads_or(stop != 0);
        if (warp == 0) st.unsat[lane] = 0;
        if (any) {
            if (stop) {
                // frames that pass the pre-check return the raw-sign word; every other exit returns hard(Qv)
                uint32_t ok0 = it == 0 ? (stop & ~fail) : 0u;
                if (ok0) snapshot(ok0, true);
                if (stop & ~ok0) snapshot(stop & ~ok0, false);
            }
            if (warp == 0) {
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (stop >> b & 1) iters[lane * 4 + b] = (fail >> b & 1) ? -1 : it;
                st.done[lane] = done | stop;
            }
        }
        const int all = __syncthreads_and(((done | stop) & 0xfu) == 0xfu);
        if (all || last) break;

        // ---- one layered iteration, horizontal_layered.rs:105-110
        for (int l = 0; l < p.num_levels; ++l) {
            const int r0 = __ldg(p.level_ptr + l), r1 = __ldg(p.level_ptr + l + 1);
            for (int ri = r0 + warp; ri < r1; ri += kGWarps) {
                const int c = __ldg(p.level_rows + ri);
                const int e0 = __ldg(g.row_ptr + c), d = __ldg(g.row_ptr + c + 1) - e0;
                if (d == 0) continue;
#define LDPC_ROW_CASE(D_) case D_: layered_tile_row<F, RULE, IS_I8, HLIM, D_, Q, R>(qv, rcv, g.col_idx, e0, d, lane, tb); break;
                switch (d) {
                    LDPC_ROW_CASE(1) LDPC_ROW_CASE(2) LDPC_ROW_CASE(3) LDPC_ROW_CASE(4) LDPC_ROW_CASE(5) LDPC_ROW_CASE(6)
                    LDPC_ROW_CASE(7) LDPC_ROW_CASE(8) LDPC_ROW_CASE(9) LDPC_ROW_CASE(10)
                    default: layered_tile_row<F, RULE, IS_I8, HLIM, 0, Q, R>(qv, rcv, g.col_idx, e0, d, lane, tb); break;
                }
#undef LDPC_ROW_CASE
            }
            __syncthreads();
        }
    }
}

}  // namespace

// ---- launchers ---------------------------------------------------------------------------------
template <class F>
static bool launch_flood_float_t(const GenericLaunch& L, cudaStream_t stream) {
    FloodFloatParams<F> p;
    p.g = L.graph; p.msg = static_cast<F*>(L.msg); p.hbit = L.hbit; p.in = static_cast<const F*>(L.in);
    p.raw0 = L.raw0; p.final_hard = L.final_hard; p.iters = L.iters; p.max_iter = L.max_iter;
    dim3 grid((unsigned)L.num_tiles), block(kGWarps * 32);
    switch (L.rule) {
        case kPhi: flood_float_kernel<F, kPhi><<<grid, block, 0, stream>>>(p); break;
        case kTanh: flood_float_kernel<F, kTanh><<<grid, block, 0, stream>>>(p); break;
        case kMinstarapprox: flood_float_kernel<F, kMinstarapprox><<<grid, block, 0, stream>>>(p); break;
        default: flood_float_kernel<F, kAminstar><<<grid, block, 0, stream>>>(p); break;
    }
    LDPC_CUDA_CHECK(cudaGetLastError());
    return true;
}

bool launch_flood_float(const GenericLaunch& L, cudaStream_t stream) {
    return L.is_f64 ? launch_flood_float_t<double>(L, stream) : launch_flood_float_t<float>(L, stream);
}

template <class F, bool IS_I8>
static bool launch_layered_t(const GenericLaunch& L, cudaStream_t stream) {
    using Q = typename std::conditional<IS_I8, int16_t, F>::type;
    using R = typename std::conditional<IS_I8, int8_t, F>::type;
    LayeredParams<Q, R> p;
    p.g = L.graph; p.level_ptr = L.level_ptr; p.level_rows = L.level_rows; p.num_levels = L.num_levels;
    p.qv = static_cast<Q*>(L.in_out_q); p.rcv = static_cast<R*>(L.msg); p.raw0 = L.raw0; p.final_hard = L.final_hard;
    p.iters = L.iters; p.max_iter = L.max_iter;
    dim3 grid((unsigned)L.num_tiles), block(kGWarps * 32);
    if (IS_I8) {
        if (L.rule == kMinstarapprox) {
            if (L.hardlimit) layered_kernel<F, kMinstarapprox, IS_I8, true><<<grid, block, 0, stream>>>(p);
            else layered_kernel<F, kMinstarapprox, IS_I8, false><<<grid, block, 0, stream>>>(p);
        } else {
            if (L.hardlimit) layered_kernel<F, kAminstar, IS_I8, true><<<grid, block, 0, stream>>>(p);
            else layered_kernel<F, kAminstar, IS_I8, false><<<grid, block, 0, stream>>>(p);
        }
    } else {
        switch (L.rule) {
            case kPhi: layered_kernel<F, kPhi, IS_I8, false><<<grid, block, 0, stream>>>(p); break;
            case kTanh: layered_kernel<F, kTanh, IS_I8, false><<<grid, block, 0, stream>>>(p); break;
            case kMinstarapprox: layered_kernel<F, kMinstarapprox, IS_I8, false><<<grid, block, 0, stream>>>(p); break;
            default: layered_kernel<F, kAminstar, IS_I8, false><<<grid, block, 0, stream>>>(p); break;
        }
    }
    LDPC_CUDA_CHECK(cudaGetLastError());
    return true;
}

bool launch_layered(const GenericLaunch& L, cudaStream_t stream) {
    if (L.is_i8) return launch_layered_t<float, true>(L, stream);
    return L.is_f64 ? launch_layered_t<double, false>(L, stream) : launch_layered_t<float, false>(L, stream);
}

int generic_max_row_degree() { return kRuleMaxD; }

}  // namespace ldpc
