// ldpc_toolbox_b200/csrc/flood_float_impl.cuh — K2: flooding-schedule BP with the float rules, one
// thread-block cluster per 128-frame tile.
//
//   replaces flooding::Decoder<A>::decode for A in {Phi, Tanh, Minstarapprox, Aminstar} x {f64, f32}
//   reference src/decoder/flooding.rs:51-125, src/decoder/arithmetic.rs:140-156 (variable node),
//   check rules in rules.cuh (arithmetic.rs:214-246, :347-379, :487-521, :942-999)
//
// Layout: 128-frame tiles, frame index fastest ([node][128] values); a lane owns 4 consecutive
// frames and runs the reference's per-frame arithmetic on them one after the other, in the
// reference's order (sum order, product order, first argmin).  Messages live in one array in
// row-major edge order, overwritten in place by each pass, like K1 (flood_i8.cu); hard decisions
// are kept per edge so the syndrome of iteration i is read during the check pass of iteration i+1.
// The CTAs of the tile's cluster split every pass (see flood_float_kernel).
// Included by one translation unit per float type (flood_float_f32.cu, flood_float_f64.cu).
#pragma once
#include "bp_common.cuh"

namespace ldpc {
namespace {

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
    F* post;                // test hook (may be null): [tiles][n][128] output_llrs of each frame's last processed iteration (flooding.rs:111-125)
};

// L2-only access to the message state: with a thread-block cluster per tile the check pass and the
// variable pass of one edge may run on different SMs, whose L1 caches are not coherent.
template <class T> __device__ __forceinline__ V4<T> ld4cg(const T* base, size_t node, int lane);
template <> __device__ __forceinline__ V4<float> ld4cg(const float* base, size_t node, int lane) {
    float4 t = __ldcg(reinterpret_cast<const float4*>(base + node * kTileFrames + lane * 4));
    return {{t.x, t.y, t.z, t.w}};
}
template <> __device__ __forceinline__ V4<double> ld4cg(const double* base, size_t node, int lane) {
    const double2* p = reinterpret_cast<const double2*>(base + node * kTileFrames + lane * 4);
    double2 a = __ldcg(p), b = __ldcg(p + 1);
    return {{a.x, a.y, b.x, b.y}};
}
template <class T> __device__ __forceinline__ void st4cg(T* base, size_t node, int lane, const V4<T>& x);
template <> __device__ __forceinline__ void st4cg(float* base, size_t node, int lane, const V4<float>& x) {
    __stcg(reinterpret_cast<float4*>(base + node * kTileFrames + lane * 4), make_float4(x.v[0], x.v[1], x.v[2], x.v[3]));
}
template <> __device__ __forceinline__ void st4cg(double* base, size_t node, int lane, const V4<double>& x) {
    double2* p = reinterpret_cast<double2*>(base + node * kTileFrames + lane * 4);
    __stcg(p, make_double2(x.v[0], x.v[1]));
    __stcg(p + 1, make_double2(x.v[2], x.v[3]));
}

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
    for (int j = 0; j < d; ++j) xs[j] = ld4cg<F>(msg, e0 + j, lane);
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
    for (int j = 0; j < d; ++j) st4cg<F>(msg, e0 + j, lane, xs[j]);
}

// One variable node (arithmetic.rs:140-156): llr = input + sum of the check messages in cols[v] order,
// message on edge e = llr - c_e.  DT > 0: all lines are requested before the first is used.
template <class F, int DT>
__device__ __forceinline__ void flood_var_node(F* __restrict__ msg, uint8_t* __restrict__ hbit, const V4<F>& inp,
                                               const int* __restrict__ col_edge, int d_rt, int lane, F* __restrict__ post, size_t v,
                                               uint32_t live) {
    constexpr int CAP = DT > 0 ? DT : 1;
    const int d = DT > 0 ? DT : d_rt;
    F sum[4] = {F(0), F(0), F(0), F(0)};
    int e[CAP];
    V4<F> c[CAP];
    if (DT > 0) {
#pragma unroll
        for (int j = 0; j < d; ++j) e[j] = __ldg(col_edge + j);
#pragma unroll
        for (int j = 0; j < d; ++j) c[j] = ld4cg<F>(msg, (size_t)e[j], lane);
#pragma unroll
        for (int j = 0; j < d; ++j)
#pragma unroll
            for (int f = 0; f < 4; ++f) sum[f] += c[j].v[f];
    } else {
        for (int j = 0; j < d; ++j) {
            V4<F> t = ld4cg<F>(msg, (size_t)__ldg(col_edge + j), lane);
#pragma unroll
            for (int f = 0; f < 4; ++f) sum[f] += t.v[f];
        }
    }
    F llr[4];
    uint32_t hb = 0;
#pragma unroll
    for (int f = 0; f < 4; ++f) { llr[f] = inp.v[f] + sum[f]; hb |= (uint32_t)(llr[f] <= F(0)) << f; }
    if (post) {                       // frames that already stopped keep the posterior of their last iteration
        V4<F> pw = ld4cg<F>(post, v, lane);
#pragma unroll
        for (int f = 0; f < 4; ++f)
            if (live >> f & 1) pw.v[f] = llr[f];
        st4cg<F>(post, v, lane, pw);
    }
    if (DT > 0) {
#pragma unroll
        for (int j = 0; j < d; ++j) {
#pragma unroll
            for (int f = 0; f < 4; ++f) c[j].v[f] = llr[f] - c[j].v[f];
            st4cg<F>(msg, (size_t)e[j], lane, c[j]);
            __stcg(hbit + (size_t)e[j] * kLanes + lane, (uint8_t)hb);
        }
    } else {
        for (int j = 0; j < d; ++j) {
            const size_t ee = (size_t)__ldg(col_edge + j);
            V4<F> t = ld4cg<F>(msg, ee, lane);
#pragma unroll
            for (int f = 0; f < 4; ++f) t.v[f] = llr[f] - t.v[f];
            st4cg<F>(msg, ee, lane, t);
            __stcg(hbit + ee * kLanes + lane, (uint8_t)hb);
        }
    }
}

// One thread-block cluster owns a tile for the whole decode: its CTAs split the checks (check pass)
// and the variables (variable pass), a cluster barrier separates the passes, and the per-frame stop
// decision is taken identically in every CTA from the OR of all CTAs' syndrome words, read through
// distributed shared memory.  Cluster size 1 is the plain one-CTA-per-tile case (large batches).
// two CTAs per SM for f32 (128 registers; with one CTA the same code is 2.4x slower on 5G-NR BG1), one for f64
template <class F, int RULE>
__global__ void __launch_bounds__(kGWarps * 32, sizeof(F) == 8 ? 1 : 2) flood_float_kernel(FloodFloatParams<F> p) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    __shared__ uint32_t s_unsat[2][kLanes];      // double-buffered by iteration parity: no reset race across CTAs
    __shared__ uint32_t s_done[kLanes];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = rank * kGWarps + warp, nw = C * kGWarps;      // this warp among the cluster's warps
    const size_t tile = blockIdx.x / (unsigned)C;
    const DeviceGraph& g = p.g;
    F* msg = p.msg + tile * (size_t)g.E * kTileFrames;
    uint8_t* hbit = p.hbit + tile * (size_t)g.E * kLanes;
    const F* in = p.in + tile * (size_t)g.n * kTileFrames;
    const uint8_t* raw0 = p.raw0 + tile * (size_t)g.n * kLanes;
    uint8_t* fin = p.final_hard + tile * (size_t)g.n * kLanes;
    int32_t* iters = p.iters + tile * kTileFrames;
    F* post = p.post ? p.post + tile * (size_t)g.n * kTileFrames : nullptr;
    if (threadIdx.x < kLanes) { s_unsat[0][threadIdx.x] = 0; s_unsat[1][threadIdx.x] = 0; s_done[threadIdx.x] = 0; }

    // flooding.rs:88-100
    for (int v = gw; v < g.n; v += nw) {
        V4<F> w = ld4<F>(in, (size_t)v, lane);
        uint8_t hb = raw0[(size_t)v * kLanes + lane];
        for (int q = __ldg(g.col_ptr + v); q < __ldg(g.col_ptr + v + 1); ++q) {
            size_t e = (size_t)__ldg(g.col_edge + q);
            st4cg<F>(msg, e, lane, w);
            __stcg(hbit + e * kLanes + lane, hb);
        }
    }
    cluster.sync();

    for (int it = 1;; ++it) {
        const bool last = it > p.max_iter;
        const int buf = it & 1;
        uint32_t synd = 0;
        for (int c = gw; c < g.m; c += nw) {                             // flooding.rs:102-109
            const int e0 = __ldg(g.row_ptr + c), d = __ldg(g.row_ptr + c + 1) - e0;
            uint32_t hb = 0;
            for (int j = 0; j < d; ++j) hb ^= __ldcg(hbit + (size_t)(e0 + j) * kLanes + lane);
            synd |= hb;
            if (last || d == 0) continue;
            // O(d) rules are unrolled up to degree 20 (5G-NR base graph 1 has rows of degree 19), the
            // O(d^2) transcendental fold of Min*-approx up to 10
            constexpr int kUnrollMax = (rule_is_minstar(RULE) || sizeof(F) == 8) ? 10 : 20;     // (f64: code size / build time)
#define LDPC_CHK_CASE(D_) case D_: flood_check_row<F, RULE, (D_ <= kUnrollMax ? D_ : 0)>(msg, (size_t)e0, d, lane); break;
            switch (d) {
                LDPC_CHK_CASE(1) LDPC_CHK_CASE(2) LDPC_CHK_CASE(3) LDPC_CHK_CASE(4) LDPC_CHK_CASE(5) LDPC_CHK_CASE(6)
                LDPC_CHK_CASE(7) LDPC_CHK_CASE(8) LDPC_CHK_CASE(9) LDPC_CHK_CASE(10) LDPC_CHK_CASE(11) LDPC_CHK_CASE(12)
                LDPC_CHK_CASE(13) LDPC_CHK_CASE(14) LDPC_CHK_CASE(15) LDPC_CHK_CASE(16) LDPC_CHK_CASE(17) LDPC_CHK_CASE(18)
                LDPC_CHK_CASE(19) LDPC_CHK_CASE(20)
                default: flood_check_row<F, RULE, 0>(msg, (size_t)e0, d, lane); break;
            }
#undef LDPC_CHK_CASE
        }
        if (synd) atomicOr(&s_unsat[buf][lane], synd);
        cluster.sync();
        uint32_t unsat = 0;
        for (int r = 0; r < C; ++r) unsat |= *cluster.map_shared_rank(&s_unsat[buf][lane], r);
        const uint32_t done = s_done[lane];
        uint32_t stop = ~unsat & ~done & 0xfu, fail = 0;
        if (last) { fail = unsat & ~done & 0xfu; stop |= fail; }
        const int any = __syncthreads_or(stop != 0);
        if (warp == 0) s_unsat[buf ^ 1][lane] = 0;       // next iteration's buffer; its last readers passed the barrier above
        if (any) {
            if (stop) {
                for (int v = gw; v < g.n; v += nw) {
                    size_t o = (size_t)v * kLanes + lane;
                    int p0 = __ldg(g.col_ptr + v), p1 = __ldg(g.col_ptr + v + 1);
                    uint32_t hb;
                    if (p1 > p0) hb = __ldcg(hbit + (size_t)__ldg(g.col_edge + p0) * kLanes + lane);
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
                if (rank == 0) {
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (stop >> b & 1) iters[lane * 4 + b] = (fail >> b & 1) ? -1 : it - 1;
                }
                s_done[lane] = done | stop;
            }
        }
        const int all = __syncthreads_and(((done | stop) & 0xfu) == 0xfu);
        if (all || last) break;

        const uint32_t live = ~s_done[lane] & 0xfu;
        for (int v = gw; v < g.n; v += nw) {                             // flooding.rs:111-125
            const int p0 = __ldg(g.col_ptr + v), d = __ldg(g.col_ptr + v + 1) - p0;
            const V4<F> inp = ld4<F>(in, (size_t)v, lane);
            const int* ce = g.col_edge + p0;
#define LDPC_VAR_CASE(D_) case D_: flood_var_node<F, D_>(msg, hbit, inp, ce, d, lane, post, (size_t)v, live); break;
            switch (d) {
                case 0: break;
                LDPC_VAR_CASE(1) LDPC_VAR_CASE(2) LDPC_VAR_CASE(3) LDPC_VAR_CASE(4) LDPC_VAR_CASE(5) LDPC_VAR_CASE(6)
                LDPC_VAR_CASE(7) LDPC_VAR_CASE(8) LDPC_VAR_CASE(9) LDPC_VAR_CASE(10) LDPC_VAR_CASE(11) LDPC_VAR_CASE(12)
                LDPC_VAR_CASE(13)
                default: flood_var_node<F, 0>(msg, hbit, inp, ce, d, lane, post, (size_t)v, live); break;
            }
#undef LDPC_VAR_CASE
        }
        cluster.sync();
    }
    cluster.sync();          // no CTA may leave while a peer can still read its shared memory
}


template <class F>
static bool launch_flood_float_t(const GenericLaunch& L, cudaStream_t stream) {
    FloodFloatParams<F> p;
    p.g = L.graph; p.msg = static_cast<F*>(L.msg); p.hbit = L.hbit; p.in = static_cast<const F*>(L.in);
    p.raw0 = L.raw0; p.final_hard = L.final_hard; p.iters = L.iters; p.max_iter = L.max_iter;
    p.post = static_cast<F*>(L.post);
    const int C = L.cluster >= 1 && L.cluster <= 8 ? L.cluster : 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)L.num_tiles * (unsigned)C);
    cfg.blockDim = dim3(kGWarps * 32);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    switch (L.rule) {
        case kPhi: LDPC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, flood_float_kernel<F, kPhi>, p)); break;
        case kTanh: LDPC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, flood_float_kernel<F, kTanh>, p)); break;
        case kMinstarapprox: LDPC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, flood_float_kernel<F, kMinstarapprox>, p)); break;
        // bit-exact libm mode: instantiated for f32 only (f64 evaluates the same expression either way)
        case kMinstarapproxExact: LDPC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, flood_float_kernel<F, (sizeof(F) == 4 ? kMinstarapproxExact : kMinstarapprox)>, p)); break;
        case kAminstarExact: LDPC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, flood_float_kernel<F, (sizeof(F) == 4 ? kAminstarExact : kAminstar)>, p)); break;
        default: LDPC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, flood_float_kernel<F, kAminstar>, p)); break;
    }
    LDPC_CUDA_CHECK(cudaGetLastError());
    return true;
}


}  // namespace
}  // namespace ldpc
