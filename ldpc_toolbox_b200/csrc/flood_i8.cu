// ldpc_toolbox_b200/csrc/flood_i8.cu — K1: flooding-schedule BP with the reference's 8-bit
// arithmetics (16 variants), bit-exact.
//
// Replaces, for a whole tile of 128 frames at a time:
//   flooding::Decoder::decode            reference src/decoder/flooding.rs:51-125
//   Minstarapproxi8*::send_check_messages reference src/decoder/arithmetic.rs:718-754
//   Aminstari8*::send_check_messages      reference src/decoder/arithmetic.rs:1130-1192
//   impl_send_var_messages_i8             reference src/decoder/arithmetic.rs:622-654
//   check_llrs / hard_decisions           reference src/decoder.rs:157-174
//
// One CTA owns one tile for the whole decode (all iterations); its warps split the check nodes
// (check pass) and the variable nodes (variable pass).  Messages live in ONE array msg[E][32]
// (uint32 = 4 frames) in row-major edge order: the check pass reads v->c values and overwrites
// them in place with c->v values, the variable pass does the reverse.  Every access is a
// 128-byte line.  The syndrome of iteration i is evaluated during the check pass of iteration
// i+1 from a compact hard-decision plane (4 bits per lane), so early termination costs no extra
// pass over the messages; iteration 0 is the reference's pre-check on the raw LLR signs.
//
// Exactness notes: the min* fold g(a,b)=max(0,min(a,b)-T[|a-b|]) is not
// associative, so for every excluded edge j the others are folded left-to-right in row order;
// the only sharing is the common prefix fold(x_0..x_{j-1}).  T is a 128-entry table in shared
// memory (exactly one 4-byte word per bank => conflict-free for any index pattern).
#include <string>

#include "decoder_impl.hpp"
#include "device_common.cuh"

namespace ldpc {

namespace {

struct FloodI8Params {
    DeviceGraph g;
    uint32_t* msg;          // [tiles][E][32]
    const uint32_t* inq;    // [tiles][n][32]   quantised channel LLRs (int8 x4)
    uint8_t* hard;          // [tiles][n][32]   4 hard bits per lane; iteration 0 = raw LLR signs
    uint8_t* final_hard;    // [tiles][n][32]   snapshot taken when a frame stops
    int32_t* iters;         // [tiles*128]      iterations, or -1 on failure
    int max_iter;
    int jones, deg1clip;
};

constexpr int kWarps = 8;            // warps per CTA
constexpr int kMaxUnrollD = 10;      // check degrees with a fully unrolled register path
constexpr int kMaxGenericD = 64;     // larger rows are rejected when the decoder is built

// g(a, acc) of arithmetic.rs:741 on non-negative ints; negT[t] = -T[t]
__device__ __forceinline__ int gop(int a, int acc, const int8_t* __restrict__ negT) {
    int mn = min(a, acc);
    int t = a + acc - 2 * mn;                       // |a - acc|
    return __viaddmax_s32_relu(mn, (int)negT[t], 0);  // max(mn - T[t], 0)
}

// h(a, acc) of arithmetic.rs:1155-1157
__device__ __forceinline__ int hop(int a, int acc, const int8_t* __restrict__ negT) {
    int mn = min(a, acc);
    int t = a + acc - 2 * mn;
    int s = min(a + acc, 127);                      // i8 saturating_add
    return max(mn + (int)negT[t] - (int)negT[s], 0);
}

__device__ __forceinline__ int hardlimit(int mag) { return mag >= 100 ? 127 : mag; }   // arithmetic.rs:812-824

// magnitude word (4 x 0..127) and sign word (bit 7 of each byte) -> 4 x int8 two's complement,
// with -0 = 0
__device__ __forceinline__ uint32_t apply_signs(uint32_t mag, uint32_t sgn) {
    uint32_t m1 = sgn >> 7;
    uint32_t m7 = sgn - m1;                          // 0x7f in negative bytes
    return ((mag ^ m7) + m1) ^ sgn;
}

template <int D, bool HLIM>
__device__ __forceinline__ void check_minstar(uint32_t* __restrict__ mrow, const int8_t* __restrict__ negT) {
    uint32_t x[D];
#pragma unroll
    for (int j = 0; j < D; ++j) x[j] = ld_stream(mrow + j * kLanes);
    uint32_t S = 0;
#pragma unroll
    for (int j = 0; j < D; ++j) S ^= x[j];
    uint32_t om[D];
#pragma unroll
    for (int j = 0; j < D; ++j) om[j] = 0;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        int a[D], r[D];
#pragma unroll
        for (int j = 0; j < D; ++j) a[j] = abs((int)(int8_t)(x[j] >> (8 * f)));
        if (D == 2) {
            r[0] = a[1];
            r[1] = a[0];
        } else {
            int acc = a[1];
#pragma unroll
            for (int i = 2; i < D; ++i) acc = gop(a[i], acc, negT);
            r[0] = acc;
            int P = a[0];                            // fold(x_0 .. x_{j-1})
#pragma unroll
            for (int j = 1; j < D; ++j) {
                acc = P;
#pragma unroll
                for (int i = j + 1; i < D; ++i) acc = gop(a[i], acc, negT);
                r[j] = acc;
                if (j < D - 1) P = gop(a[j], P, negT);
            }
        }
#pragma unroll
        for (int j = 0; j < D; ++j) {
            int mg = HLIM ? hardlimit(r[j]) : r[j];
            om[j] |= (uint32_t)mg << (8 * f);
        }
    }
#pragma unroll
    for (int j = 0; j < D; ++j) st_stream(mrow + j * kLanes, apply_signs(om[j], (S ^ x[j]) & 0x80808080u));
}

template <int D, bool HLIM>
__device__ __forceinline__ void check_aminstar(uint32_t* __restrict__ mrow, const int8_t* __restrict__ negT) {
    uint32_t x[D];
#pragma unroll
    for (int j = 0; j < D; ++j) x[j] = ld_stream(mrow + j * kLanes);
    uint32_t S = 0;
#pragma unroll
    for (int j = 0; j < D; ++j) S ^= x[j];
    uint32_t om[D];
#pragma unroll
    for (int j = 0; j < D; ++j) om[j] = 0;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        int a[D];
#pragma unroll
        for (int j = 0; j < D; ++j) a[j] = abs((int)(int8_t)(x[j] >> (8 * f)));
        int amin = a[0], arg = 0;                    // first minimum (min_by_key)
#pragma unroll
        for (int j = 1; j < D; ++j)
            if (a[j] < amin) { amin = a[j]; arg = j; }
        int delta = -1;
#pragma unroll
        for (int j = 0; j < D; ++j) {
            if (j != arg) delta = delta < 0 ? a[j] : hop(a[j], delta, negT);
        }
        int d2 = hop(delta, amin, negT);
        if (HLIM) { delta = hardlimit(delta); d2 = hardlimit(d2); }
#pragma unroll
        for (int j = 0; j < D; ++j) om[j] |= (uint32_t)(j == arg ? delta : d2) << (8 * f);
    }
#pragma unroll
    for (int j = 0; j < D; ++j) st_stream(mrow + j * kLanes, apply_signs(om[j], (S ^ x[j]) & 0x80808080u));
}

// any degree up to kMaxGenericD; inputs staged in local memory
template <bool AMIN, bool HLIM>
__device__ __noinline__ void check_generic(uint32_t* __restrict__ mrow, int d, const int8_t* __restrict__ negT) {
    uint32_t x[kMaxGenericD], om[kMaxGenericD];
    uint32_t S = 0;
    for (int j = 0; j < d; ++j) { x[j] = ld_stream(mrow + j * kLanes); S ^= x[j]; om[j] = 0; }
    for (int f = 0; f < 4; ++f) {
        if (AMIN) {
            int amin = 1 << 20, arg = 0;
            for (int j = 0; j < d; ++j) {
                int a = abs((int)(int8_t)(x[j] >> (8 * f)));
                if (a < amin) { amin = a; arg = j; }
            }
            int delta = -1;
            for (int j = 0; j < d; ++j) {
                if (j == arg) continue;
                int a = abs((int)(int8_t)(x[j] >> (8 * f)));
                delta = delta < 0 ? a : hop(a, delta, negT);
            }
            int d2 = hop(delta, amin, negT);
            if (HLIM) { delta = hardlimit(delta); d2 = hardlimit(d2); }
            for (int j = 0; j < d; ++j) om[j] |= (uint32_t)(j == arg ? delta : d2) << (8 * f);
        } else {
            int P = 0;
            for (int j = 0; j < d; ++j) {
                int acc = -1;
                if (j > 0) acc = P;
                for (int i = j + 1; i < d; ++i) {
                    int a = abs((int)(int8_t)(x[i] >> (8 * f)));
                    acc = acc < 0 ? a : gop(a, acc, negT);
                }
                int mg = HLIM ? hardlimit(acc) : acc;
                om[j] |= (uint32_t)mg << (8 * f);
                int aj = abs((int)(int8_t)(x[j] >> (8 * f)));
                P = j == 0 ? aj : gop(aj, P, negT);
            }
        }
    }
    for (int j = 0; j < d; ++j) st_stream(mrow + j * kLanes, apply_signs(om[j], (S ^ x[j]) & 0x80808080u));
}

template <bool AMIN, bool HLIM, int D>
__device__ __forceinline__ void check_fixed(uint32_t* mrow, const int8_t* negT) {
    if (AMIN) check_aminstar<D, HLIM>(mrow, negT);
    else check_minstar<D, HLIM>(mrow, negT);
}

template <bool AMIN, bool HLIM>
__device__ __forceinline__ void check_dispatch(uint32_t* mrow, int d, const int8_t* negT) {
    switch (d) {
        case 0: break;
        case 1: break;   // the reference panics; such graphs are refused before launch
        case 2: check_fixed<AMIN, HLIM, 2>(mrow, negT); break;
        case 3: check_fixed<AMIN, HLIM, 3>(mrow, negT); break;
        case 4: check_fixed<AMIN, HLIM, 4>(mrow, negT); break;
        case 5: check_fixed<AMIN, HLIM, 5>(mrow, negT); break;
        case 6: check_fixed<AMIN, HLIM, 6>(mrow, negT); break;
        case 7: check_fixed<AMIN, HLIM, 7>(mrow, negT); break;
        case 8: check_fixed<AMIN, HLIM, 8>(mrow, negT); break;
        case 9: check_fixed<AMIN, HLIM, 9>(mrow, negT); break;
        case 10: check_fixed<AMIN, HLIM, 10>(mrow, negT); break;
        default: check_generic<AMIN, HLIM>(mrow, d, negT); break;
    }
}

// ---- variable node, arithmetic.rs:622-654, on 2 x (2 frames as s16x2) per lane -----------------
// Everything is kept in a biased unsigned domain (value + 128 per term) so that plain 32-bit adds
// never carry between the two 16-bit halves; the bias is removed inside the DPX add-min op.
struct VarAcc { uint32_t lo, hi; };

__device__ __forceinline__ VarAcc widen_biased(uint32_t w_i8x4) {
    uint32_t b = w_i8x4 ^ 0x80808080u;              // int8 + 128, per byte
    return {prmt(b, 0, 0x4140), prmt(b, 0, 0x4342)};
}

__device__ __forceinline__ uint32_t rep16(int v) { return ((uint32_t)v & 0xffffu) * 0x00010001u; }

__device__ __forceinline__ uint32_t clip127(uint32_t v) {      // per-half clamp to [-127, 127]
    return __vmaxs2(__vmins2(v, 0x007f007fu), 0xff81ff81u);
}

// Finishes a variable node once the biased sum (input + all check messages) is known.
// Returns the 4 hard bits; out(j, word) is called with the int8x4 message for slot j.
template <class GetMsg, class PutMsg>
__device__ __forceinline__ uint32_t var_finish(VarAcc sum, int d, bool jones, GetMsg get, PutMsg put) {
    // true L = sum - 128*(d+1)
    uint32_t Llo = __vadd2(sum.lo, rep16(-128 * (d + 1)));
    uint32_t Lhi = __vadd2(sum.hi, rep16(-128 * (d + 1)));
    uint32_t base_lo, base_hi, negK;
    if (jones) {                                      // arithmetic.rs:806-810: L = clip(L)
        Llo = clip127(Llo);
        Lhi = clip127(Lhi);
        base_lo = __vadd2(Llo, rep16(384));           // L + 128 + 256 >= 257 > any biased message
        base_hi = __vadd2(Lhi, rep16(384));
        negK = rep16(-256);
    } else {
        base_lo = sum.lo;
        base_hi = sum.hi;
        negK = rep16(-128 * d);
    }
    for (int j = 0; j < d; ++j) {
        VarAcc c = widen_biased(get(j));
        // clip(L - c_j) = clamp((base - c'_j) - K, -127, 127); base >= c'_j in both halves
        uint32_t vlo = __vmaxs2(__viaddmin_s16x2(base_lo - c.lo, negK, 0x007f007fu), 0xff81ff81u);
        uint32_t vhi = __vmaxs2(__viaddmin_s16x2(base_hi - c.hi, negK, 0x007f007fu), 0xff81ff81u);
        put(j, prmt(vlo, vhi, 0x6420));
    }
    // hard decision L <= 0  <=>  sign bit of (L - 1); clip() never changes it
    uint32_t zlo = __vadd2(Llo, 0xffffffffu), zhi = __vadd2(Lhi, 0xffffffffu);
    uint32_t sb = prmt(zlo, zhi, 0x7531);             // high bytes of the four halves
    return pack_bits4((sb >> 7) & 0x01010101u);
}

template <int D>
__device__ __forceinline__ uint32_t var_fixed(uint32_t* __restrict__ msg, const int* __restrict__ ce, uint32_t inw,
                                              bool jones, bool deg1clip, int lane) {
    uint32_t w[D];
    int e[D];
#pragma unroll
    for (int j = 0; j < D; ++j) e[j] = __ldg(ce + j);
#pragma unroll
    for (int j = 0; j < D; ++j) w[j] = ld_stream(msg + (size_t)e[j] * kLanes + lane);
    VarAcc sum = widen_biased(inw);
    if (D == 1 && deg1clip) {                         // arithmetic.rs:826-842, biased: [12, 244]
        sum.lo = __vmaxu2(__vminu2(sum.lo, rep16(244)), rep16(12));
        sum.hi = __vmaxu2(__vminu2(sum.hi, rep16(244)), rep16(12));
    }
#pragma unroll
    for (int j = 0; j < D; ++j) {
        VarAcc c = widen_biased(w[j]);
        sum.lo += c.lo;
        sum.hi += c.hi;
    }
    return var_finish(sum, D, jones,
                      [&](int j) { return w[j]; },
                      [&](int j, uint32_t v) { st_stream(msg + (size_t)e[j] * kLanes + lane, v); });
}

__device__ __noinline__ uint32_t var_generic(uint32_t* __restrict__ msg, const int* __restrict__ ce, int d, uint32_t inw,
                                             bool jones, int lane) {
    VarAcc sum = widen_biased(inw);
    for (int j = 0; j < d; ++j) {
        VarAcc c = widen_biased(ld_stream(msg + (size_t)__ldg(ce + j) * kLanes + lane));
        sum.lo += c.lo;
        sum.hi += c.hi;
    }
    return var_finish(sum, d, jones,
                      [&](int j) { return ld_stream(msg + (size_t)__ldg(ce + j) * kLanes + lane); },
                      [&](int j, uint32_t v) { st_stream(msg + (size_t)__ldg(ce + j) * kLanes + lane, v); });
}

template <bool AMIN, bool HLIM>
__global__ void __launch_bounds__(kWarps * 32) flood_i8_kernel(FloodI8Params p) {
    __shared__ __align__(128) int8_t negT[128];
    __shared__ uint32_t s_unsat[kLanes];
    __shared__ uint32_t s_done[kLanes];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t tile = blockIdx.x;
    const DeviceGraph& g = p.g;
    uint32_t* msg = p.msg + tile * (size_t)g.E * kLanes;
    const uint32_t* inq = p.inq + tile * (size_t)g.n * kLanes;
    uint8_t* hard = p.hard + tile * (size_t)g.n * kLanes;
    uint8_t* fin = p.final_hard + tile * (size_t)g.n * kLanes;
    int32_t* iters = p.iters + tile * kTileFrames;

    if (threadIdx.x < 128) {
        // T[t] = round(8 ln(1 + e^{-t/8})) = #{theta in {1,3,5,9,13,22} : t < theta}  
        int t = threadIdx.x;
        negT[t] = (int8_t)-((t < 1) + (t < 3) + (t < 5) + (t < 9) + (t < 13) + (t < 22));
    }
    if (threadIdx.x < kLanes) { s_unsat[threadIdx.x] = 0; s_done[threadIdx.x] = 0; }

    // flooding.rs:88-100: first variable messages are the quantised channel LLRs
    for (int v = warp; v < g.n; v += kWarps) {
        uint32_t w = __ldg(inq + (size_t)v * kLanes + lane);
        int p0 = __ldg(g.col_ptr + v), p1 = __ldg(g.col_ptr + v + 1);
        for (int q = p0; q < p1; ++q) st_stream(msg + (size_t)__ldg(g.col_edge + q) * kLanes + lane, w);
    }
    __syncthreads();

    for (int it = 1;; ++it) {
        const bool last = it > p.max_iter;           // only the syndrome of iteration max_iter is left
        uint32_t synd = 0;
        for (int c = warp; c < g.m; c += kWarps) {
            int e0 = __ldg(g.row_ptr + c), d = __ldg(g.row_ptr + c + 1) - e0;
            uint32_t hb = 0;
            for (int j = 0; j < d; ++j) hb ^= hard[(size_t)__ldg(g.col_idx + e0 + j) * kLanes + lane];
            synd |= hb;
            if (!last) check_dispatch<AMIN, HLIM>(msg + (size_t)e0 * kLanes + lane, d, negT);
        }
        if (synd) atomicOr(&s_unsat[lane], synd);
        __syncthreads();
        const uint32_t unsat = s_unsat[lane], done = s_done[lane];
        // frames whose hard decisions of iteration it-1 satisfy every check stop now
        // (flooding.rs:57-64 for it-1 == 0, :69-79 otherwise)
        uint32_t stop = ~unsat & ~done & 0xfu;
        uint32_t fail = 0;
        if (last) { fail = unsat & ~done & 0xfu; stop |= fail; }      // flooding.rs:81-85
        const int any = __syncthreads_or(stop != 0);
        if (warp == 0) s_unsat[lane] = 0;
        if (any) {
            if (stop) {
                for (int v = warp; v < g.n; v += kWarps) {
                    size_t o = (size_t)v * kLanes + lane;
                    fin[o] = (uint8_t)((fin[o] & ~stop) | (hard[o] & stop));
                }
            }
            if (warp == 0) {
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (stop >> b & 1) iters[lane * 4 + b] = (fail >> b & 1) ? -1 : it - 1;
                s_done[lane] = done | stop;
            }
        }
        const int all = __syncthreads_and(((done | stop) & 0xfu) == 0xfu);
        if (all || last) break;

        for (int v = warp; v < g.n; v += kWarps) {
            int p0 = __ldg(g.col_ptr + v), d = __ldg(g.col_ptr + v + 1) - p0;
            uint32_t inw = __ldg(inq + (size_t)v * kLanes + lane);
            const int* ce = g.col_edge + p0;
            uint32_t hb;
            switch (d) {
                case 0: {   // no checks: L = input (degree-one clip does not apply)
                    VarAcc s = widen_biased(inw);
                    hb = var_finish(s, 0, p.jones != 0, [&](int) { return 0u; }, [&](int, uint32_t) {});
                    break;
                }
                case 1: hb = var_fixed<1>(msg, ce, inw, p.jones != 0, p.deg1clip != 0, lane); break;
                case 2: hb = var_fixed<2>(msg, ce, inw, p.jones != 0, false, lane); break;
                case 3: hb = var_fixed<3>(msg, ce, inw, p.jones != 0, false, lane); break;
                case 4: hb = var_fixed<4>(msg, ce, inw, p.jones != 0, false, lane); break;
                case 5: hb = var_fixed<5>(msg, ce, inw, p.jones != 0, false, lane); break;
                case 6: hb = var_fixed<6>(msg, ce, inw, p.jones != 0, false, lane); break;
                case 7: hb = var_fixed<7>(msg, ce, inw, p.jones != 0, false, lane); break;
                case 8: hb = var_fixed<8>(msg, ce, inw, p.jones != 0, false, lane); break;
                default: hb = var_generic(msg, ce, d, inw, p.jones != 0, lane); break;
            }
            hard[(size_t)v * kLanes + lane] = (uint8_t)hb;
        }
        __syncthreads();
    }
}

}  // namespace

bool launch_flood_i8(const FloodI8Launch& L, cudaStream_t stream) {
    FloodI8Params p;
    p.g = L.graph;
    p.msg = L.msg; p.inq = L.inq; p.hard = L.hard; p.final_hard = L.final_hard; p.iters = L.iters;
    p.max_iter = L.max_iter; p.jones = L.jones; p.deg1clip = L.deg1clip;
    dim3 grid((unsigned)L.num_tiles), block(kWarps * 32);
    if (L.aminstar) {
        if (L.hardlimit) flood_i8_kernel<true, true><<<grid, block, 0, stream>>>(p);
        else flood_i8_kernel<true, false><<<grid, block, 0, stream>>>(p);
    } else {
        if (L.hardlimit) flood_i8_kernel<false, true><<<grid, block, 0, stream>>>(p);
        else flood_i8_kernel<false, false><<<grid, block, 0, stream>>>(p);
    }
    LDPC_CUDA_CHECK(cudaGetLastError());
    return true;
}

int flood_i8_max_row_degree() { return kMaxGenericD; }

}  // namespace ldpc
