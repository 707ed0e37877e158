// ldpc_toolbox_b200/csrc/flood_i8.cu — K1: flooding-schedule BP with the reference's 8-bit
// arithmetics (16 variants), bit-exact.
//
// Replaces, for a whole tile of frames at a time:
//   flooding::Decoder::decode             reference src/decoder/flooding.rs:51-125
//   Minstarapproxi8*::send_check_messages reference src/decoder/arithmetic.rs:718-754
//   Aminstari8*::send_check_messages      reference src/decoder/arithmetic.rs:1130-1192
//   impl_send_var_messages_i8             reference src/decoder/arithmetic.rs:622-654
//   check_llrs / hard_decisions           reference src/decoder.rs:157-174
//
// One CTA owns one tile for the whole decode (all iterations); its warps split the check nodes
// (check pass) and the variable nodes (variable pass).  A tile is 32 lanes x NW words x 4 frames
// (NW = 1: 128 frames, 128-byte lines; NW = 4: 512 frames, 512-byte lines moved with 128-bit
// loads — random 512-byte granules stream HBM at ~6-7 TB/s where 128-byte ones stop near
// 3.6-4 TB/s, tools/membench.cu).  Messages live in ONE array msg[E][32][NW] (uint32 = 4 frames)
// in row-major edge order: the check pass reads v->c values and overwrites them in place with
// c->v values, the variable pass does the reverse.  Hard decisions are kept per EDGE
// (hbit[E][32], 4*NW bits per lane) so a check reads the bits of its variables from the same
// contiguous span as its messages: the syndrome of iteration i is evaluated during the check pass
// of iteration i+1 with no gather and no extra pass; iteration 0 is the reference's pre-check on
// the raw LLR signs.
//
// Exactness notes: the min* fold g(a,b)=max(0,min(a,b)-T[|a-b|]) is not
// associative, so for every excluded edge j the others are folded left-to-right in row order;
// the only sharing is the common prefix fold(x_0..x_{j-1}).  With d = a - acc,
//      g(a, acc) = max(0, acc + U[d]),   U[d] = min(d, 0) - T[|d|].
// The unrolled path carries a chain as the index of its next table read, idx = a_next - acc: the
// table returns V[idx] = idx - U[idx] and the following index is min(V[idx] + (a' - a), a'), so a
// fold step is one LDS.U8 and one VIADDMNMX (check_word).  U, V and T are small tables in shared
// memory (the A-Min* variants and the generic-degree path use U and T directly).
//
// Storage format: every int8 quantity in HBM (messages, channel LLRs) is kept in offset binary,
// byte = value + 128 in [1, 255].  |v| of four frames is then ONE instruction (VABSDIFF4.U8 against
// 0x80808080), the variable node's carry-free SWAR sums need no re-biasing, and signs travel in bit 7.
//
// Pipe balance: the check pass is bound by the table reads (one shared-memory wavefront per warp
// read) together with the half-rate integer pipe (VIADDMNMX, PRMT, LOP3), while the FMA pipe (IMAD)
// idles.  Plain adds / subtracts / shift-and-ors are therefore written as mad.lo with a multiplier
// taken from a kernel parameter (ptxas cannot fold it), which pins them to the FMA pipe.
// Check inputs are staged per warp in shared memory by TMA bulk copies (see the check pass).
#include <cooperative_groups.h>

#include <cstdlib>
#include <string>

#include "decoder_impl.hpp"
#include "device_common.cuh"

namespace ldpc {

namespace {

struct FloodI8Params {
    DeviceGraph g;
    VarClasses vc;
    uint32_t* msg;          // [tiles][E][32][NW]
    void* hbit;             // [tiles][E][32]     hard decision of the edge's variable, 4*NW bits per lane
    const uint32_t* inq;    // [tiles][n][32][NW] quantised channel LLRs (int8 x4 per word)
    const void* raw0;       // [tiles][n][32]     raw-sign hard decisions (x <= 0.0), 4*NW bits per lane
    void* final_hard;       // [tiles][n][32]     snapshot taken when a frame stops
    const RowMeta* row_meta;  // [m]              per-row record: first edge, degree, staircase-fusion flags (decoder_impl.hpp)
    const int* snap_src;      // [n]              where a variable's hard decisions live (see the stop logic)
    int snap_n;               //                  variables whose final decisions are read back (the caller's output_len)
    void* cbit;             // [tiles][2][m][32]  hard decisions of the fused variables by iteration parity, 4*NW bits per lane
    int chunk_rows;         // rows per chunk dealt to a warp in the check pass (power of two)
    int fuse_var_off;       // the variable fused between rows r-1 and r is r + fuse_var_off (staircase: k - 1)
    int max_row_deg;        // largest check degree of the code
    int32_t* iters;         // [tiles*128*NW]     iterations, or -1 on failure
    int max_iter;
    int num_tiles;
    int jones, deg1clip;
    // opaque multipliers for FMA-pipe integer arithmetic (see header): -1, 1, -2, 255, 2^8, 2^16, 2^24
    int c_m1, c_one, c_m2, c_ff, c_sh8, c_sh16, c_sh24;
};


struct Consts { int m1, one, m2, ff, sh[4]; };

#ifndef LDPC_I8_WARPS
#define LDPC_I8_WARPS 8
#endif
#ifndef LDPC_I8_MINBLOCKS
#define LDPC_I8_MINBLOCKS 2
#endif
#ifndef LDPC_I8_GROUPS
#define LDPC_I8_GROUPS 1
#endif
constexpr int kWarps = LDPC_I8_WARPS;            // warps per tile (one warp group)
// Warp groups per CTA.  2: a CTA owns two tiles and runs them half an iteration apart, in lockstep —
// while group 0 is in its check pass (table reads, integer pipe) group 1 is in its variable pass
// (HBM gathers) and vice versa, so the two passes overlap instead of competing with themselves.
constexpr int kGroups = LDPC_I8_GROUPS;
constexpr int kCtaThreads = kWarps * 32 * kGroups;

#ifdef LDPC_I8_PROFILE
// experiment-only: SM-clock cycles thread 0 of every CTA spent in {init, check pass, stop logic, variable pass}, [4] = CTA-iterations
__device__ unsigned long long g_i8_prof[8];
#define PROF_T(var) long long var = clock64()
#define PROF_ADD(slot, a, b) do { if (threadIdx.x == 0) atomicAdd(&g_i8_prof[slot], (unsigned long long)((b) - (a))); } while (0)
#else
#define PROF_T(var) do {} while (0)
#define PROF_ADD(slot, a, b) do {} while (0)
#endif
constexpr int kMaxGenericD = 64;     // larger rows are rejected when the decoder is built
constexpr int align128(int x) { return (x + 127) & ~127; }

__device__ __forceinline__ int lane_of_thread() { return threadIdx.x & 31; }
// barriers of one warp group (named barrier 1 + group, kWarps * 32 threads); with one group per CTA they
// are the plain CTA barriers
__device__ __forceinline__ void group_sync(int grp) {
    if (kGroups == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(kWarps * 32) : "memory");
}
__device__ __forceinline__ int group_or(int grp, int pred) {
    if (kGroups == 1) return __syncthreads_or(pred);
    int r;
    asm volatile("{\n.reg .pred p, q;\nsetp.ne.s32 q, %1, 0;\nbar.red.or.pred p, %2, %3, q;\nselp.s32 %0, 1, 0, p;\n}"
                 : "=r"(r) : "r"(pred), "r"(grp + 1), "r"(kWarps * 32) : "memory");
    return r;
}
__device__ __forceinline__ int group_and(int grp, int pred) {
    if (kGroups == 1) return __syncthreads_and(pred);
    int r;
    asm volatile("{\n.reg .pred p, q;\nsetp.ne.s32 q, %1, 0;\nbar.red.and.pred p, %2, %3, q;\nselp.s32 %0, 1, 0, p;\n}"
                 : "=r"(r) : "r"(pred), "r"(grp + 1), "r"(kWarps * 32) : "memory");
    return r;
}

template <int NW> struct Lane { uint32_t w[NW]; };
template <int NW> struct HBitsT { using type = uint8_t; };
template <> struct HBitsT<4> { using type = uint16_t; };

template <int NW>
__device__ __forceinline__ Lane<NW> ld_lane(const uint32_t* base, size_t node, int lane) {
    Lane<NW> r;
    const uint32_t* p = base + (node * kLanes + lane) * NW;
    if (NW == 4) {
        uint4 v = __ldcg(reinterpret_cast<const uint4*>(p));
        r.w[0] = v.x; r.w[1 % NW] = v.y; r.w[2 % NW] = v.z; r.w[3 % NW] = v.w;
    } else {
        r.w[0] = __ldcg(p);
    }
    return r;
}
template <int NW>
__device__ __forceinline__ void st_lane(uint32_t* base, size_t node, int lane, const Lane<NW>& v) {
    uint32_t* p = base + (node * kLanes + lane) * NW;
    if (NW == 4) __stcg(reinterpret_cast<uint4*>(p), make_uint4(v.w[0], v.w[1 % NW], v.w[2 % NW], v.w[3 % NW]));
    else __stcg(p, v.w[0]);
}

// ---- TMA bulk copy global -> shared, completion on an mbarrier (one elected lane issues it) ------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct Tables {
    int8_t U[256];      // U[d + 127], d = a - acc in [-127, 127]
    uint8_t V[256];     // V[d + 127] = d - U[d] = max(d, 0) + T[|d|]
    int8_t Tp[128];     // T[t]
};

__device__ __forceinline__ int table_T(int t) {
    // T[t] = round(8 ln(1 + e^{-t/8})) = #{theta in {1,3,5,9,13,22} : t < theta}  
    return (t < 1) + (t < 3) + (t < 5) + (t < 9) + (t < 13) + (t < 22);
}

// a*b + c on the FMA pipe (b must come from a kernel parameter so that it stays an IMAD)
__device__ __forceinline__ int imad(int a, int b, int c) {
    int r;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// g(a, acc) of arithmetic.rs:741 on non-negative ints: IMAD (a - acc), LDS, VIADDMNMX.RELU
__device__ __forceinline__ int gop(int a, int acc, const Tables& tb, const Consts& k) {
    return __viaddmax_s32_relu(acc, (int)tb.U[imad(acc, k.m1, a) + 127], 0);
}

// h(a, acc) of arithmetic.rs:1155-1157
__device__ __forceinline__ int hop(int a, int acc, const Tables& tb, const Consts& k) {
    int s = min(a + acc, 127);                      // i8 saturating_add
    return max(acc + (int)tb.U[imad(acc, k.m1, a) + 127] + (int)tb.Tp[s], 0);
}

__device__ __forceinline__ int hardlimit(int mag) { return mag >= 100 ? 127 : mag; }   // arithmetic.rs:812-824

// magnitude word (4 x 0..127) and sign word (0x80 in negative bytes) -> offset-binary bytes 128 +- mag
__device__ __forceinline__ uint32_t apply_signs(uint32_t mag, uint32_t sgn, const Consts& k) {
    uint32_t negmask = (uint32_t)imad((int)(sgn >> 7), k.ff, 0);        // 0xff in negative bytes
    uint32_t t = mag & negmask;
    return (uint32_t)imad((int)t, k.m2, (int)(mag + 0x80808080u));      // per byte: 128 + mag - 2*t, no borrows
}

// |v| of the four offset-binary bytes of a word
__device__ __forceinline__ uint32_t abs4(uint32_t x) { return __vabsdiffu4(x, 0x80808080u); }
__device__ __forceinline__ int byte_of(uint32_t x, int f) { return (int)__byte_perm(x, 0, 0x4440 + f); }

// sign word of output j: XOR of the signs of all OTHER inputs.  Sx = XOR of the D offset-binary words
// (bit 7 set = non-negative), so bit 7 of Sx ^ x_j is parity(D-1) ^ XOR_{i != j} neg_i.
template <int DPARITY>
__device__ __forceinline__ uint32_t sign_excluding(uint32_t Sx, uint32_t xj) {
    return (DPARITY & 1) ? (~(Sx ^ xj) & 0x80808080u) : ((Sx ^ xj) & 0x80808080u);   // D-1 odd <=> D even
}
__device__ __forceinline__ uint32_t sign_excluding_rt(uint32_t Sx, uint32_t xj, int d) {
    uint32_t s = (Sx ^ xj) & 0x80808080u;
    return (d & 1) ? s : s ^ 0x80808080u;
}

// One word (4 frames) of a degree-D check: x[0..D) in (offset binary), c->v words out (in place).
// Frame slots whose bit is set in `skip` (every lane's frame there has stopped) are not computed.
// SKIP = false is the steady-state path: no per-frame branches, so the four frames of the word form
// one basic block and their fold chains interleave (the chains are latency-bound: IMAD -> LDS -> VIADDMNMX).
template <int D, bool AMIN, bool HLIM, bool SKIP>
__device__ __forceinline__ void check_word(uint32_t (&x)[D], uint32_t skip, const Tables& tb, const Consts& k) {
    uint32_t S = 0, A[D], om[D];
#pragma unroll
    for (int j = 0; j < D; ++j) { S ^= x[j]; A[j] = abs4(x[j]); om[j] = 0; }
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        if (SKIP && (skip >> f & 1)) continue;
        int a[D], r[D];
#pragma unroll
        for (int j = 0; j < D; ++j) a[j] = byte_of(A[j], f);
        if (AMIN) {
            int amin = a[0], arg = 0;                // first minimum (min_by_key)
#pragma unroll
            for (int j = 1; j < D; ++j)
                if (a[j] < amin) { amin = a[j]; arg = j; }
            int delta = -1;
#pragma unroll
            for (int j = 0; j < D; ++j)
                if (j != arg) delta = delta < 0 ? a[j] : hop(a[j], delta, tb, k);
            int d2 = hop(delta, amin, tb, k);
#pragma unroll
            for (int j = 0; j < D; ++j) r[j] = j == arg ? delta : d2;
        } else if (D == 2) {
            r[0] = a[1];
            r[1] = a[0];
        } else {
            // Difference form of the same folds: a chain is carried as idx = (next input) - acc, the
            // table returns V[idx] = idx - U[idx], and the index of the following step is
            //   a' - max(acc + U[idx], 0) = min(V[idx] + (a' - a), a')
            // so a fold step is LDS + VIADDMNMX (the subtraction a' - a is shared by all chains), and a
            // chain ends with acc = max(a - V[idx], 0).
            int d1[D - 1];
#pragma unroll
            for (int i = 0; i + 1 < D; ++i) d1[i] = imad(a[i], k.m1, a[i + 1]);
            // fold inputs a[i0..D) into a chain whose state is idx = a[i0] - acc
            auto run = [&](int idx, int i0) {
                int out = 0;
#pragma unroll
                for (int i = 2; i < D; ++i) {
                    if (i < i0) continue;
                    const int v = (int)tb.V[idx + 127];
                    if (i + 1 < D) idx = __viaddmin_s32(v, d1[i], a[i + 1]);
                    else out = __viaddmax_s32_relu(a[i], -v, 0);
                }
                return out;
            };
            r[0] = run(d1[1], 2);                                   // acc = a1, then a2 ..
            r[1] = run(imad(a[0], k.m1, a[2]), 2);                  // acc = a0, then a2 ..
            int dP = d1[0];                                         // a_{j-1} - P_{j-1}, P_1 = a0
#pragma unroll
            for (int j = 2; j < D; ++j) {                           // this lookup evaluates P_j = g(a_{j-1}, P_{j-1})
                const int v = (int)tb.V[dP + 127];
                if (j == D - 1) {
                    r[j] = __viaddmax_s32_relu(a[j - 1], -v, 0);    // output D-1 is P_{D-1} itself
                } else {
                    dP = __viaddmin_s32(v, d1[j - 1], a[j]);        // a_j - P_j
                    r[j] = run(imad(dP, k.one, d1[j]), j + 1);      // a_{j+1} - P_j, then a_{j+2} ..
                }
            }
        }
#pragma unroll
        for (int j = 0; j < D; ++j) {
            int mg = HLIM ? hardlimit(r[j]) : r[j];
            om[j] = f == 0 ? (uint32_t)mg : (uint32_t)imad(mg, k.sh[f], (int)om[j]);     // om |= mg << 8f on the FMA pipe
        }
    }
#pragma unroll
    for (int j = 0; j < D; ++j) x[j] = apply_signs(om[j], sign_excluding<D - 1>(S, x[j]), k);
}

// any degree up to kMaxGenericD; one word at a time, inputs staged in local memory
template <int NW, bool AMIN, bool HLIM>
__device__ __noinline__ void check_generic(uint32_t* __restrict__ msg, size_t e0, int d, int lane, uint32_t skip, const Tables& tb,
                                           const Consts& k) {
    for (int q = 0; q < NW; ++q) {
        uint32_t x[kMaxGenericD], A[kMaxGenericD], om[kMaxGenericD];
        uint32_t S = 0;
        uint32_t* mrow = msg + (e0 * kLanes + lane) * NW + q;
        for (int j = 0; j < d; ++j) { x[j] = __ldcg(mrow + (size_t)j * kLanes * NW); S ^= x[j]; A[j] = abs4(x[j]); om[j] = 0; }
        for (int f = 0; f < 4; ++f) {
            if (skip >> (4 * q + f) & 1) continue;
            if (AMIN) {
                int amin = 1 << 20, arg = 0;
                for (int j = 0; j < d; ++j) {
                    int a = (int)((A[j] >> (8 * f)) & 0xffu);
                    if (a < amin) { amin = a; arg = j; }
                }
                int delta = -1;
                for (int j = 0; j < d; ++j) {
                    if (j == arg) continue;
                    int a = (int)((A[j] >> (8 * f)) & 0xffu);
                    delta = delta < 0 ? a : hop(a, delta, tb, k);
                }
                int d2 = hop(delta, amin, tb, k);
                if (HLIM) { delta = hardlimit(delta); d2 = hardlimit(d2); }
                for (int j = 0; j < d; ++j) om[j] |= (uint32_t)(j == arg ? delta : d2) << (8 * f);
            } else {
                int P = 0;
                for (int j = 0; j < d; ++j) {
                    int acc = -1;
                    if (j > 0) acc = P;
                    for (int i = j + 1; i < d; ++i) {
                        int a = (int)((A[i] >> (8 * f)) & 0xffu);
                        acc = acc < 0 ? a : gop(a, acc, tb, k);
                    }
                    int mg = HLIM ? hardlimit(acc) : acc;
                    om[j] |= (uint32_t)mg << (8 * f);
                    int aj = (int)((A[j] >> (8 * f)) & 0xffu);
                    P = j == 0 ? aj : gop(aj, P, tb, k);
                }
            }
        }
        for (int j = 0; j < d; ++j) __stcg(mrow + (size_t)j * kLanes * NW, apply_signs(om[j], sign_excluding_rt(S, x[j], d), k));
    }
}

// Rows of degree MAXD < d <= the stage capacity of the wide kernels (DVB-S2 rates >= 3/5: row degrees 10 ... 30).
// The row's message lines stay in the warp's shared-memory stage: a word (4 frames) at a time, the four frames
// fold side by side (four independent chains hide the IMAD -> LDS -> VIADDMNMX latency), operands are read from
// the stage and every output word is written back in place as soon as its chain is done — input j is not needed
// any more then (later chains start from the prefix P_{j+1}, which already contains it).  Same order as
// check_word: for every excluded j the others are folded left to right, sharing only the prefix (arithmetic.rs:722-751).
template <int NW, bool AMIN, bool HLIM>
__device__ __noinline__ void check_wide(uint32_t* __restrict__ sx, int d, int lane, uint32_t skip, const Tables& tb, const Consts& k) {
    constexpr int L = kLanes * NW;                       // words between consecutive lines of the stage
    if (AMIN) {
        for (int q = 0; q < NW; ++q) {
            if (((skip >> (4 * q)) & 0xfu) == 0xfu) continue;
            uint32_t* w = sx + lane * NW + q;            // this lane's word of line j is w[j * L]
            uint32_t S = 0;
            for (int j = 0; j < d; ++j) S ^= w[j * L];
            // first minimum per frame, then the O(d) fold of the others (arithmetic.rs:1130-1192)
            int amin[4], arg[4], delta[4], d2[4];
#pragma unroll
            for (int f = 0; f < 4; ++f) { amin[f] = 1 << 20; arg[f] = 0; delta[f] = -1; }
            for (int j = 0; j < d; ++j) {
                const uint32_t A = abs4(w[j * L]);
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    const int a = byte_of(A, f);
                    if (a < amin[f]) { amin[f] = a; arg[f] = j; }
                }
            }
            for (int j = 0; j < d; ++j) {
                const uint32_t A = abs4(w[j * L]);
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    if (j == arg[f]) continue;
                    const int a = byte_of(A, f);
                    delta[f] = delta[f] < 0 ? a : hop(a, delta[f], tb, k);
                }
            }
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                d2[f] = hop(delta[f], amin[f], tb, k);
                if (HLIM) { delta[f] = hardlimit(delta[f]); d2[f] = hardlimit(d2[f]); }
            }
            for (int j = 0; j < d; ++j) {
                uint32_t om = 0;
#pragma unroll
                for (int f = 0; f < 4; ++f) om |= (uint32_t)(j == arg[f] ? delta[f] : d2[f]) << (8 * f);
                const uint32_t xw = w[j * L];
                w[j * L] = apply_signs(om, sign_excluding_rt(S, xw, d), k);
            }
        }
    } else {
        // W words (4 W frames) fold side by side: 8 independent chains on 512-frame tiles
        constexpr int W = NW >= 2 ? 2 : 1;
        for (int q = 0; q < NW; q += W) {
            if (((skip >> (4 * q)) & ((1u << (4 * W)) - 1u)) == ((1u << (4 * W)) - 1u)) continue;
            uint32_t* w = sx + lane * NW + q;            // word u of this lane in line j is w[j * L + u]
            uint32_t S[W];
#pragma unroll
            for (int u = 0; u < W; ++u) S[u] = 0;
            for (int j = 0; j < d; ++j)
#pragma unroll
                for (int u = 0; u < W; ++u) S[u] ^= w[j * L + u];
            int P[W][4];
            for (int j = 0; j < d; ++j) {
                int acc[W][4];
                int i = j + 1;
                if (j == 0) {                             // the chain of output 0 starts from input 1
#pragma unroll
                    for (int u = 0; u < W; ++u) {
                        const uint32_t A = abs4(w[L + u]);
#pragma unroll
                        for (int f = 0; f < 4; ++f) acc[u][f] = byte_of(A, f);
                    }
                    i = 2;
                } else {
#pragma unroll
                    for (int u = 0; u < W; ++u)
#pragma unroll
                        for (int f = 0; f < 4; ++f) acc[u][f] = P[u][f];
                }
                uint32_t nxt[W];                          // operands of the next step are in flight during this one
#pragma unroll
                for (int u = 0; u < W; ++u) nxt[u] = i < d ? w[i * L + u] : 0u;
                for (; i < d; ++i) {
                    uint32_t A[W];
#pragma unroll
                    for (int u = 0; u < W; ++u) A[u] = abs4(nxt[u]);
                    if (i + 1 < d) {
#pragma unroll
                        for (int u = 0; u < W; ++u) nxt[u] = w[(i + 1) * L + u];
                    }
#pragma unroll
                    for (int u = 0; u < W; ++u)
#pragma unroll
                        for (int f = 0; f < 4; ++f) acc[u][f] = gop(byte_of(A[u], f), acc[u][f], tb, k);
                }
#pragma unroll
                for (int u = 0; u < W; ++u) {
                    const uint32_t xw = w[j * L + u];
                    const uint32_t Aj = abs4(xw);
                    uint32_t om = 0;
#pragma unroll
                    for (int f = 0; f < 4; ++f) {
                        const int mg = HLIM ? hardlimit(acc[u][f]) : acc[u][f];
                        om |= (uint32_t)mg << (8 * f);
                        const int aj = byte_of(Aj, f);
                        P[u][f] = j == 0 ? aj : gop(aj, P[u][f], tb, k);
                    }
                    w[j * L + u] = apply_signs(om, sign_excluding_rt(S[u], xw, d), k);
                }
            }
        }
    }
}

// ---- variable node, arithmetic.rs:622-654, on 2 x (2 frames as s16x2) per word -----------------
// Offset-binary bytes (value + 128) widen to unsigned 16-bit halves, so plain 32-bit adds never carry
// between the two halves; the bias is removed inside the DPX add-min op.
struct VarAcc { uint32_t lo, hi; };

__device__ __forceinline__ VarAcc widen(uint32_t w_ob) { return {prmt(w_ob, 0, 0x4140), prmt(w_ob, 0, 0x4342)}; }

__device__ __forceinline__ uint32_t rep16(int v) { return ((uint32_t)v & 0xffffu) * 0x00010001u; }

__device__ __forceinline__ uint32_t clip127(uint32_t v) {      // per-half clamp to [-127, 127]
    return __vmaxs2(__vmins2(v, 0x007f007fu), 0xff81ff81u);
}

struct VarConsts { uint32_t negL, negK; bool jones; };
__device__ __forceinline__ VarConsts var_consts(int d, bool jones) {
    // negK also re-adds the +128 of the offset-binary output
    return {rep16(-128 * (d + 1)), jones ? rep16(-256 + 128) : rep16(-128 * d + 128), jones};
}

// One word of a variable node once its biased sum (input + all check messages) is known: returns
// the 4 hard bits and leaves in (base_lo, base_hi) the minuend for the outgoing messages.
__device__ __forceinline__ uint32_t var_posterior(VarAcc sum, const VarConsts& k, uint32_t& base_lo, uint32_t& base_hi) {
    uint32_t Llo = __vadd2(sum.lo, k.negL), Lhi = __vadd2(sum.hi, k.negL);   // true L = sum - 128*(d+1)
    if (k.jones) {                                    // arithmetic.rs:806-810: L = clip(L)
        Llo = clip127(Llo);
        Lhi = clip127(Lhi);
        base_lo = __vadd2(Llo, rep16(384));           // L + 128 + 256 >= 257 > any biased message
        base_hi = __vadd2(Lhi, rep16(384));
    } else {
        base_lo = sum.lo;
        base_hi = sum.hi;
    }
    // hard decision L <= 0  <=>  sign bit of (L - 1); clip() never changes it
    uint32_t zlo = __vadd2(Llo, 0xffffffffu), zhi = __vadd2(Lhi, 0xffffffffu);
    return pack_bits4((prmt(zlo, zhi, 0x7531) >> 7) & 0x01010101u);
}

// offset-binary clip(L - c_j) + 128 = clamp((base - c'_j) - K + 128, 1, 255); base >= c'_j in both halves
__device__ __forceinline__ uint32_t var_message(uint32_t c_ob, uint32_t base_lo, uint32_t base_hi, uint32_t negK, const Consts& k) {
    VarAcc c = widen(c_ob);
    uint32_t ulo = (uint32_t)imad((int)c.lo, k.m1, (int)base_lo), uhi = (uint32_t)imad((int)c.hi, k.m1, (int)base_hi);
    uint32_t vlo = __vmaxs2(__viaddmin_s16x2(ulo, negK, 0x00ff00ffu), 0x00010001u);
    uint32_t vhi = __vmaxs2(__viaddmin_s16x2(uhi, negK, 0x00ff00ffu), 0x00010001u);
    return prmt(vlo, vhi, 0x6420);
}

// One check of degree D on the register path: the check-node rule word by word, then — staircase fusion,
// decoder_impl.hpp — the degree-2 variable this row shares with the previous row (slot D-2) is updated on the
// spot from `carry` (the previous row's message to it, kept in registers), x[D-2] and its channel LLRs `inl`:
// both of its outgoing messages and its hard decision are final for this iteration.  The previous row's last
// line is stored now (deferred by one row); this row's last line is kept in `carry` when the next row fuses.
template <int NW, int MAXD, int D, bool AMIN, bool HLIM>
__device__ __forceinline__ void check_fixed(Lane<NW> (&x)[MAXD], uint32_t* __restrict__ msg, size_t e0, int lane, uint32_t skip,
                                            const Tables& tb, const Consts& k, bool fuse_prev, bool fuse_next, Lane<NW>& carry,
                                            const Lane<NW>* __restrict__ s_inl, bool jones, typename HBitsT<NW>::type* __restrict__ cnew,
                                            int row, Lane<NW>* __restrict__ cslot) {
    if (skip == 0) {
#pragma unroll
        for (int q = 0; q < NW; ++q) {
            uint32_t xw[D];
#pragma unroll
            for (int j = 0; j < D; ++j) xw[j] = x[j].w[q];
            check_word<D, AMIN, HLIM, false>(xw, 0, tb, k);
#pragma unroll
            for (int j = 0; j < D; ++j) x[j].w[q] = xw[j];
        }
    } else {
#pragma unroll
        for (int q = 0; q < NW; ++q) {
            if (((skip >> (4 * q)) & 0xfu) == 0xfu) continue;
            uint32_t xw[D];
#pragma unroll
            for (int j = 0; j < D; ++j) xw[j] = x[j].w[q];
            check_word<D, AMIN, HLIM, true>(xw, skip >> (4 * q), tb, k);
#pragma unroll
            for (int j = 0; j < D; ++j) x[j].w[q] = xw[j];
        }
    }
    if (fuse_prev) {
        // everything the fused update needs is fetched here, after the fold: nothing of it is live across check_word
        const VarConsts vk = var_consts(2, jones);
        uint32_t hb = 0;
        const Lane<NW> inl = s_inl[lane];
#ifdef LDPC_I8_CARRY_SMEM
        carry = cslot[lane];
#endif
#pragma unroll
        for (int q = 0; q < NW; ++q) {
            VarAcc sum = widen(inl.w[q]);
            const VarAcc ca = widen(carry.w[q]), cb = widen(x[D - 2].w[q]);
            sum.lo = (uint32_t)imad((int)ca.lo, k.one, (int)sum.lo) + cb.lo;
            sum.hi = (uint32_t)imad((int)ca.hi, k.one, (int)sum.hi) + cb.hi;
            uint32_t blo, bhi;
            hb |= var_posterior(sum, vk, blo, bhi) << (4 * q);
            carry.w[q] = var_message(carry.w[q], blo, bhi, vk.negK, k);
            x[D - 2].w[q] = var_message(x[D - 2].w[q], blo, bhi, vk.negK, k);
        }
        st_lane<NW>(msg, e0 - 1, lane, carry);            // the previous row's last edge is this row's first edge - 1
        cnew[(size_t)(row - 1) * kLanes + lane] = (typename HBitsT<NW>::type)hb;
    }
#pragma unroll
    for (int j = 0; j + 1 < D; ++j) st_lane<NW>(msg, e0 + j, lane, x[j]);
#ifdef LDPC_I8_CARRY_SMEM
    if (fuse_next) cslot[lane] = x[D - 1];
#else
    if (fuse_next) carry = x[D - 1];
#endif
    else st_lane<NW>(msg, e0 + D - 1, lane, x[D - 1]);
}

// U variables of degree D per warp iteration: all index loads, then all message loads, then math.
// (A shared-memory cp.async pipeline like the check pass's was measured slower here: 45.5 vs 35 ms per
// iteration of the bench workload — the gather is latency-bound per line, not issue-bound.)
template <int NW, int D, int U>
__device__ __forceinline__ void var_class(uint32_t* __restrict__ msg, typename HBitsT<NW>::type* __restrict__ hbit,
                                          const uint32_t* __restrict__ inq, const int* __restrict__ vlist,
                                          const int* __restrict__ vedges, int count, bool jones, bool deg1clip,
                                          uint32_t skip, int warp, int nwarps, int lane, const Consts& kc) {
    const VarConsts k = var_consts(D, jones);
    // indices of this warp's next group are fetched while the current group's lines are in flight
    auto load_idx = [&](int i, int (&e)[U][D], int (&v)[U]) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int ii = i + u < count ? i + u : i;
            v[u] = __ldg(vlist + ii);
#pragma unroll
            for (int j = 0; j < D; ++j) e[u][j] = __ldg(vedges + (size_t)ii * D + j);
        }
    };
    int i = warp * U;
    if (i >= count) return;
    int e[U][D], v[U];
    load_idx(i, e, v);
    for (; i < count; i += nwarps * U) {
        Lane<NW> w[U][D], inw[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            ok[u] = i + u < count;
            inw[u] = ld_lane<NW>(inq, (size_t)v[u], lane);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < D; ++j) w[u][j] = ld_lane<NW>(msg, (size_t)e[u][j], lane);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            uint32_t hb = 0;
#pragma unroll
            for (int q = 0; q < NW; ++q) {
                if (skip != 0 && ((skip >> (4 * q)) & 0xfu) == 0xfu) continue;
                VarAcc sum = widen(inw[u].w[q]);
                if (D == 1 && deg1clip) {             // arithmetic.rs:826-842, biased: [12, 244]
                    sum.lo = __vmaxu2(__vminu2(sum.lo, rep16(244)), rep16(12));
                    sum.hi = __vmaxu2(__vminu2(sum.hi, rep16(244)), rep16(12));
                }
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    VarAcc c = widen(w[u][j].w[q]);
                    sum.lo = (uint32_t)imad((int)c.lo, kc.one, (int)sum.lo);
                    sum.hi = (uint32_t)imad((int)c.hi, kc.one, (int)sum.hi);
                }
                uint32_t blo, bhi;
                hb |= var_posterior(sum, k, blo, bhi) << (4 * q);
#pragma unroll
                for (int j = 0; j < D; ++j) w[u][j].w[q] = var_message(w[u][j].w[q], blo, bhi, k.negK, kc);
            }
#pragma unroll
            for (int j = 0; j < D; ++j) {
                st_lane<NW>(msg, (size_t)e[u][j], lane, w[u][j]);
                hbit[(size_t)e[u][j] * kLanes + lane] = (typename HBitsT<NW>::type)hb;
            }
        }
        if (i + nwarps * U < count) load_idx(i + nwarps * U, e, v);
    }
}

// variables of any degree > 8 (listed in vlist; edges through col_ptr / col_edge)
template <int NW>
__device__ __noinline__ void var_generic_class(uint32_t* __restrict__ msg, typename HBitsT<NW>::type* __restrict__ hbit,
                                               const uint32_t* __restrict__ inq, const DeviceGraph& g,
                                               const int* __restrict__ vlist, int count, bool jones, int warp, int nwarps,
                                               int lane, const Consts& kc) {
    for (int i = warp; i < count; i += nwarps) {
        int v = __ldg(vlist + i);
        int p0 = __ldg(g.col_ptr + v), d = __ldg(g.col_ptr + v + 1) - p0;
        const int* ce = g.col_edge + p0;
        const VarConsts k = var_consts(d, jones);
        uint32_t hb = 0;
        for (int q = 0; q < NW; ++q) {
            VarAcc sum = widen(__ldg(inq + ((size_t)v * kLanes + lane) * NW + q));
            for (int j = 0; j < d; ++j) {
                VarAcc c = widen(__ldcg(msg + ((size_t)__ldg(ce + j) * kLanes + lane) * NW + q));
                sum.lo += c.lo;
                sum.hi += c.hi;
            }
            uint32_t blo, bhi;
            hb |= var_posterior(sum, k, blo, bhi) << (4 * q);
            for (int j = 0; j < d; ++j) {
                uint32_t* pm = msg + ((size_t)__ldg(ce + j) * kLanes + lane) * NW + q;
                __stcg(pm, var_message(__ldcg(pm), blo, bhi, k.negK, kc));
            }
        }
        for (int j = 0; j < d; ++j) hbit[(size_t)__ldg(ce + j) * kLanes + lane] = (typename HBitsT<NW>::type)hb;
    }
}

// WCAP = 0: the narrow kernel — two stages per warp (the next check streams in while the current one is computed),
// a stage holds a check of up to MAXD edges.  WCAP = 16 / 32: the wide kernels for codes with rows of up to WCAP
// edges — ONE stage of WCAP lines per warp (a wide row is ~10^4 instructions of folding per lane, the exposed load
// latency is a few per cent), rows above MAXD are folded from shared memory by check_wide.
template <int NW, bool AMIN, bool HLIM, int WCAP>
__global__ void __launch_bounds__(kCtaThreads, kGroups == 1 ? LDPC_I8_MINBLOCKS : 1) flood_i8_kernel(FloodI8Params p) {
    using HB = typename HBitsT<NW>::type;
    constexpr int MAXD = NW == 1 ? 10 : 8;          // check degrees with an unrolled register path
    constexpr int SCAP = WCAP ? WCAP : MAXD;        // lines a stage can hold
    constexpr int kStages = WCAP ? 1 : 2;
#ifdef LDPC_I8_GATHER_FIRST
    // experiment (kept, off): no initialisation pass — iteration 1 stages every row straight from the channel-LLR /
    // raw-sign lines of its variables, one bulk copy per line.  Saves a write and a read of the message array
    // (~16 ms of 820) but the extra code in the staging path costs the steady-state loop 7 %: 858 vs 818 ms.
    constexpr bool kGatherFirst = true;
#else
    constexpr bool kGatherFirst = false;
#endif
    constexpr uint32_t kAll = NW == 1 ? 0xfu : 0xffffu;
    constexpr int kFrames = kTileFrames * NW;
    constexpr int kMsgBytes = SCAP * kLanes * NW * 4;                 // one check's message lines
    // every region of a stage starts on a 128-byte shared-memory row: a 512-byte line read with LDS.128 (and written by
    // the TMA) then costs 4 wavefronts, not 8 — a stage size of 5184 bytes (64-byte aligned) cost 12 % of the kernel
    constexpr int kInqOff = align128(kMsgBytes + SCAP * kLanes * (int)sizeof(HB));   // channel LLRs of the variable fused with the previous row
    constexpr int kCbitOff = align128(kInqOff + kLanes * NW * 4);                    // previous hard decisions of the variable fused with the next row
#ifdef LDPC_I8_SMALL_STAGE      // experiment: the round-1 stage size (valid with LDPC_B200_FUSE=0 only)
    constexpr int kStageBytes = align128(kMsgBytes + SCAP * kLanes * (int)sizeof(HB));
#else
    constexpr int kStageBytes = align128(kCbitOff + kLanes * (int)sizeof(HB));
#endif
    extern __shared__ __align__(128) uint8_t dsm[];                    // [kGroups * kWarps][2][kStageBytes]
    __shared__ __align__(128) Tables tb;
    __shared__ uint32_t s_unsat_g[kGroups][2][kLanes];      // [iteration parity]: no reset race between the CTAs of a cluster
    __shared__ uint32_t s_done_g[kGroups][kLanes];
    __shared__ uint32_t s_skip_g[kGroups];
    __shared__ int s_fin[2][kGroups];                      // [step parity][group]: the group has finished its tile
    __shared__ __align__(8) uint64_t s_bar[kGroups * kWarps][2];      // one mbarrier per warp and stage
    if (lane_of_thread() == 0) { mbar_init(&s_bar[threadIdx.x >> 5][0], 1); mbar_init(&s_bar[threadIdx.x >> 5][1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    uint32_t bar_phase = 0;                                 // bit s = parity the next wait on stage s uses

    // A thread-block cluster of C CTAs may own the tile (small batches: C x kWarps warps split every pass, a
    // cluster barrier separates the passes, the syndrome words are OR-ed through distributed shared memory);
    // C = 1 is the plain one-CTA-per-tile case of large batches.
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int lane = threadIdx.x & 31, cta_warp = threadIdx.x >> 5;
    const int grp = cta_warp / kWarps, warp = cta_warp % kWarps;      // warp group (tile) and warp within it
    const int gw = rank * kWarps + warp, nw = C * kWarps;             // this warp among the warps working on the tile
    const size_t tile = (size_t)(blockIdx.x / (unsigned)C) * kGroups + grp;
    const bool active = tile < (size_t)p.num_tiles;
    auto pass_sync = [&]() {                                 // end of a pass of this tile
        if (C > 1) cluster.sync();
        else group_sync(grp);
    };
    uint32_t* const s_done = s_done_g[grp];
    uint32_t& s_skip = s_skip_g[grp];
    const DeviceGraph& g = p.g;
    const Consts kc = {p.c_m1, p.c_one, p.c_m2, p.c_ff, {1, p.c_sh8, p.c_sh16, p.c_sh24}};
    uint32_t* msg = p.msg + tile * (size_t)g.E * kLanes * NW;
    HB* hbit = static_cast<HB*>(p.hbit) + tile * (size_t)g.E * kLanes;
    const uint32_t* inq = p.inq + tile * (size_t)g.n * kLanes * NW;
    const HB* raw0 = static_cast<const HB*>(p.raw0) + tile * (size_t)g.n * kLanes;
    HB* fin = static_cast<HB*>(p.final_hard) + tile * (size_t)g.n * kLanes;
    HB* cbit = static_cast<HB*>(p.cbit) + tile * (size_t)2 * g.m * kLanes;
    int32_t* iters = p.iters + tile * kFrames;
    const bool jones = p.jones != 0, d1c = p.deg1clip != 0;

    for (int i = threadIdx.x; i < 255; i += blockDim.x) {
        int d = i - 127;
        tb.U[i] = (int8_t)(min(d, 0) - table_T(abs(d)));
        tb.V[i] = (uint8_t)(max(d, 0) + table_T(abs(d)));
    }
    if (threadIdx.x < 128) tb.Tp[threadIdx.x] = (int8_t)table_T(threadIdx.x);
    if (warp == 0) { s_unsat_g[grp][0][lane] = 0; s_unsat_g[grp][1][lane] = 0; s_done[lane] = 0; }
    if (warp == 0 && lane == 0) s_skip = 0;
    __syncthreads();
    PROF_T(pt_init0);

    // flooding.rs:88-100: the first variable messages are the quantised channel LLRs and the "iteration 0" hard
    // decisions the raw LLR signs (flooding.rs:57).  Edge-parallel, four independent edges in flight per warp.
    // (With LDPC_I8_GATHER_FIRST only rows too wide for a stage are initialised here.)
#ifndef LDPC_I8_GATHER_FIRST
    for (int e0 = gw * 4; active && e0 < g.E; e0 += nw * 4) {
        int v[4];
        Lane<NW> w[4];
        HB hb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(g.col_idx + min(e0 + u, g.E - 1));
#pragma unroll
        for (int u = 0; u < 4; ++u) { w[u] = ld_lane<NW>(inq, (size_t)v[u], lane); hb[u] = raw0[(size_t)v[u] * kLanes + lane]; }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (e0 + u < g.E) { st_lane<NW>(msg, (size_t)(e0 + u), lane, w[u]); hbit[(size_t)(e0 + u) * kLanes + lane] = hb[u]; }
    }
    for (int r = gw; active && r < g.m; r += nw) {
        const int4 mt = __ldg(reinterpret_cast<const int4*>(p.row_meta) + r);
        if ((mt.y >> 17) & 1) cbit[(size_t)r * kLanes + lane] = raw0[(size_t)mt.w * kLanes + lane];
    }
#else
    if (p.max_row_deg > SCAP) {
        for (int r = gw; active && r < g.m; r += nw) {
            const int2 mt = __ldg(reinterpret_cast<const int2*>(p.row_meta) + 2 * (size_t)r);
            const int d = mt.y & 0xffff;
            if (d <= SCAP) continue;
            for (int j = 0; j < d; ++j) {
                const int v = __ldg(g.col_idx + mt.x + j);
                st_lane<NW>(msg, (size_t)(mt.x + j), lane, ld_lane<NW>(inq, (size_t)v, lane));
                hbit[(size_t)(mt.x + j) * kLanes + lane] = raw0[(size_t)v * kLanes + lane];
            }
        }
    }
#endif
    if (C > 1) cluster.sync();
    else __syncthreads();
    PROF_T(pt_init1);
    PROF_ADD(0, pt_init0, pt_init1);

    // A group alternates check pass (+ stop logic) and variable pass.  With two groups the CTA advances
    // in steps: group g starts at step g, and a CTA-wide barrier ends every step, so the groups stay
    // half an iteration apart for the whole decode.
    int it = 1;
    bool finished = !active, in_var = false;
    for (int step = 0;; ++step) {
      if (!finished && step >= grp) {
        if (!in_var) {
        PROF_T(pt_c0);
        const bool last = it > p.max_iter;           // only the syndrome of iteration max_iter is left
        const uint32_t skip = s_skip;                // frame slots in which every lane has stopped
        uint32_t* const s_unsat = s_unsat_g[grp][it & 1];
        uint32_t synd = 0;
        // the lines about to be fetched through the async proxy (TMA) were written through the generic proxy
        asm volatile("fence.proxy.async.global;" ::: "memory");
        {
            // Per-warp double buffer in shared memory: the messages and hard bits of the warp's next check stream
            // in while the current one is being computed, at no register cost.  A check's D message lines (and
            // its hard-bit lines) are contiguous in HBM, so one elected lane moves each with a single TMA bulk
            // copy (cp.async.bulk) that completes on the stage's mbarrier.  Rows are dealt to warps in chunks of
            // kFuseChunkRows consecutive rows (staircase fusion, decoder_impl.hpp): a stage also receives the
            // channel-LLR line of the variable fused with the previous row and the previous-iteration hard
            // decisions of the variable fused with the next row.
            uint8_t* wbuf = dsm + (size_t)cta_warp * kStages * kStageBytes;
            const HB* cold = cbit + (size_t)((it - 1) & 1) * g.m * kLanes;      // hard decisions of iteration it-1
            HB* cnew = cbit + (size_t)(it & 1) * g.m * kLanes;                  // ... of iteration it
            const int chunk = p.chunk_rows;
            auto next_row = [&](int r) { const int n1 = r + 1; return (n1 & (chunk - 1)) ? n1 : n1 + (nw - 1) * chunk; };
            // the record of the row after next is fetched one step early, so issuing a stage never waits on it
            // (first edge, degree | flags): the first half of the 16-byte row record
            auto meta_of = [&](int r) { return __ldg(reinterpret_cast<const int2*>(p.row_meta) + 2 * (size_t)min(r, g.m - 1)); };
            auto issue = [&](int stage, int r, const int2& mt) {
                const int d = mt.y & 0xffff;
                if (d <= SCAP && d > 0 && lane == 0) {
                    const bool fp = (mt.y >> 16) & 1, fn = (mt.y >> 17) & 1;
                    uint8_t* sb = wbuf + (size_t)stage * kStageBytes;
                    const uint32_t mb = last ? 0u : (uint32_t)d * kLanes * NW * 4;
                    const uint32_t nh = (uint32_t)(d - (fn ? (fp ? 2 : 1) : 0));           // trailing fused slots have no hbit line
                    const uint32_t hbytes = nh * kLanes * (uint32_t)sizeof(HB);
                    const uint32_t ib = (fp && !last) ? (uint32_t)kLanes * NW * 4 : 0u, cb = fn ? (uint32_t)kLanes * (uint32_t)sizeof(HB) : 0u;
                    mbar_expect_tx(&s_bar[cta_warp][stage], mb + hbytes + ib + cb);
                    if (kGatherFirst && it == 1) {
                        // iteration 1: a row's incoming messages ARE the channel LLRs of its variables and the decisions
                        // of "iteration 0" their raw signs — one bulk copy per variable line instead of one per row
                        constexpr uint32_t kLine = kLanes * NW * 4, kHLine = kLanes * (uint32_t)sizeof(HB);
                        int vlast = 0;
                        for (int j = 0; j < d; ++j) {
                            const int v = __ldg(g.col_idx + mt.x + j);
                            if (mb) bulk_g2s(sb + (uint32_t)j * kLine, inq + (size_t)v * kLanes * NW, kLine, &s_bar[cta_warp][stage]);
                            if ((uint32_t)j < nh) bulk_g2s(sb + kMsgBytes + (uint32_t)j * kHLine, raw0 + (size_t)v * kLanes, kHLine, &s_bar[cta_warp][stage]);
                            vlast = v;
                        }
                        if (cb) bulk_g2s(sb + kCbitOff, raw0 + (size_t)vlast * kLanes, cb, &s_bar[cta_warp][stage]);
                    } else {
                        if (mb) bulk_g2s(sb, msg + (size_t)mt.x * kLanes * NW, mb, &s_bar[cta_warp][stage]);
                        if (hbytes) bulk_g2s(sb + kMsgBytes, hbit + (size_t)mt.x * kLanes, hbytes, &s_bar[cta_warp][stage]);
                        if (cb) bulk_g2s(sb + kCbitOff, cold + (size_t)r * kLanes, cb, &s_bar[cta_warp][stage]);
                    }
                    if (ib) bulk_g2s(sb + kInqOff, inq + (size_t)(r + p.fuse_var_off) * kLanes * NW, ib, &s_bar[cta_warp][stage]);
                }
            };
            int r = gw * chunk, stage = 0;
            int2 mc = make_int2(0, 0), mn = mc;
            if (r < g.m) {
                mc = meta_of(r);
                mn = meta_of(next_row(r));
                issue(0, r, mc);
            }
            // this warp's message to the variable fused with the next row: registers, or (LDPC_I8_CARRY_SMEM) a per-warp
            // shared-memory slot, which frees four registers across the fold of the next row
            Lane<NW>* const cslot = reinterpret_cast<Lane<NW>*>(dsm + (size_t)kGroups * kWarps * kStages * kStageBytes) + (size_t)cta_warp * kLanes;
#ifdef LDPC_I8_CARRY_SMEM
            Lane<NW> carry_dummy;
#define carry carry_dummy
#else
            Lane<NW> carry;
#pragma unroll
            for (int q = 0; q < NW; ++q) carry.w[q] = 0;
#endif
            uint32_t cold_prev = 0;          // previous-iteration hard decisions of the variable fused with the previous row
            for (; r < g.m; stage ^= (kStages - 1)) {
                const int2 mt = mc;
                const int e0 = mt.x, d = mt.y & 0xffff;
#ifdef LDPC_I8_NOFUSE_STATIC       // experiment: compile the fusion out (valid with LDPC_B200_FUSE=0 only)
                const bool fuse_prev = false, fuse_next = false;
#else
                const bool fuse_prev = (mt.y >> 16) & 1, fuse_next = (mt.y >> 17) & 1;
#endif
                const int rn = next_row(r);
                mc = mn;
                if (kStages == 2 && rn < g.m) {        // two stages: the next check streams in during this one
                    issue(stage ^ 1, rn, mc);
                    mn = meta_of(next_row(rn));
                }
                if (d <= SCAP && d > 0) {
                    mbar_wait(&s_bar[cta_warp][stage], (bar_phase >> stage) & 1u);
                    bar_phase ^= 1u << stage;
                }
                uint8_t* sb = wbuf + (size_t)stage * kStageBytes;
                uint32_t hb = 0;
                if (d > SCAP) {
                    for (int j = 0; j < d; ++j) hb ^= __ldcg(hbit + (size_t)(e0 + j) * kLanes + lane);
                    if (!last) check_generic<NW, AMIN, HLIM>(msg, (size_t)e0, d, lane, skip, tb, kc);
                } else if (d > MAXD) {                 // wide row: folded in the stage, then written back line by line
                    const HB* sh = reinterpret_cast<const HB*>(sb + kMsgBytes);
                    for (int j = 0; j < d; ++j) hb ^= sh[j * kLanes + lane];
                    if (!last) {
                        check_wide<NW, AMIN, HLIM>(reinterpret_cast<uint32_t*>(sb), d, lane, skip, tb, kc);
                        const Lane<NW>* sx = reinterpret_cast<const Lane<NW>*>(sb);
                        for (int j = 0; j < d; ++j) st_lane<NW>(msg, (size_t)(e0 + j), lane, sx[j * kLanes + lane]);
                    }
                } else if (d > 0) {
                    const HB* sh = reinterpret_cast<const HB*>(sb + kMsgBytes);
#pragma unroll
                    for (int j = 0; j < MAXD; ++j)
                        if (j < d - 2 || (j == d - 2 && !fuse_prev) || (j == d - 1 && !fuse_next)) hb ^= sh[j * kLanes + lane];
                    uint32_t cold_cur = 0;
                    if (fuse_next) cold_cur = reinterpret_cast<const HB*>(sb + kCbitOff)[lane];
                    if (fuse_prev) hb ^= cold_prev;
                    hb ^= cold_cur;
                    cold_prev = cold_cur;
                    if (!last) {
                        Lane<NW> x[MAXD];
                        const Lane<NW>* sx = reinterpret_cast<const Lane<NW>*>(sb);
                        const Lane<NW>* inl = reinterpret_cast<const Lane<NW>*>(sb + kInqOff);
#pragma unroll
                        for (int j = 0; j < MAXD; ++j)
                            if (j < d) x[j] = sx[j * kLanes + lane];
#define LDPC_CHECK_CASE(D_) \
    case D_: check_fixed<NW, MAXD, (D_ <= MAXD ? D_ : 2), AMIN, HLIM>(x, msg, (size_t)e0, lane, skip, tb, kc, fuse_prev, fuse_next, carry, inl, jones, cnew, r, cslot); break;
                        switch (d) {
                            LDPC_CHECK_CASE(2) LDPC_CHECK_CASE(3) LDPC_CHECK_CASE(4) LDPC_CHECK_CASE(5) LDPC_CHECK_CASE(6)
                            LDPC_CHECK_CASE(7) LDPC_CHECK_CASE(8)
                            case 9: if (MAXD >= 9) { check_fixed<NW, MAXD, (MAXD >= 9 ? 9 : 2), AMIN, HLIM>(x, msg, (size_t)e0, lane, skip, tb, kc, false, false, carry, inl, jones, cnew, r, cslot); } break;
                            case 10: if (MAXD >= 10) { check_fixed<NW, MAXD, (MAXD >= 10 ? 10 : 2), AMIN, HLIM>(x, msg, (size_t)e0, lane, skip, tb, kc, false, false, carry, inl, jones, cnew, r, cslot); } break;
                            default: break;   // degree 1 is refused before launch (the reference panics)
                        }
#undef LDPC_CHECK_CASE
                    }
                }
                synd |= hb;
                __syncwarp();          // every lane is done with this stage before it is refilled
                if (kStages == 1 && rn < g.m) {        // one stage: refill it now
                    issue(0, rn, mc);
                    mn = meta_of(next_row(rn));
                }
                r = rn;
            }
#ifdef LDPC_I8_CARRY_SMEM
#undef carry
#endif
        }
        if (synd) atomicOr(&s_unsat[lane], synd);
        pass_sync();
        PROF_T(pt_c1);
        PROF_ADD(1, pt_c0, pt_c1);
        uint32_t unsat = s_unsat[lane];
        for (int r = 1; r < C; ++r) unsat |= *cluster.map_shared_rank(&s_unsat[lane], (unsigned)((rank + r) % C));
        const uint32_t done = s_done[lane];
        // frames whose hard decisions of iteration it-1 satisfy every check stop now
        // (flooding.rs:57-64 for it-1 == 0, :69-79 otherwise)
        uint32_t stop = ~unsat & ~done & kAll;
        uint32_t fail = 0;
        if (last) { fail = unsat & ~done & kAll; stop |= fail; }      // flooding.rs:81-85
        const int any = group_or(grp, stop != 0);
        if (warp == 0) s_unsat_g[grp][(it & 1) ^ 1][lane] = 0;     // next iteration's words; their last readers are past the barrier above
        if (any) {
            if (stop) {
                // Only the first snap_n variables are ever read back (the caller's output_len).  snap_src[v] names the
                // line holding v's hard decisions of iteration it-1: >= 0 the hbit line of its first edge, <= -2 the
                // cbit line of row -2 - x (fused variable), -1 a variable without checks.  Four variables per step so the
                // index -> line -> merge chains overlap (this loop runs whenever a frame stops: every iteration in
                // the waterfall region).
                auto snap_one = [&](int v, int src) -> uint32_t {
                    if (kGatherFirst && it == 1) return raw0[(size_t)v * kLanes + lane];      // the per-edge lines are not written yet
                    if (src >= 0) return __ldcg(hbit + (size_t)src * kLanes + lane);
                    if (src <= -2) return __ldcg(cbit + ((size_t)((it - 1) & 1) * g.m + (size_t)(-2 - src)) * kLanes + lane);
                    const size_t o = (size_t)v * kLanes + lane;
                    if (it == 1) return raw0[o];
                    uint32_t hb = 0;                      // isolated variable: posterior = quantised input
                    for (int q = 0; q < NW; ++q) {
                        const uint32_t qw = __ldg(inq + o * NW + q);
#pragma unroll
                        for (int b = 0; b < 4; ++b) hb |= (uint32_t)(((qw >> (8 * b)) & 0xffu) <= 128u) << (4 * q + b);
                    }
                    return hb;
                };
                for (int v0 = gw * 4; v0 < p.snap_n; v0 += nw * 4) {
                    int src[4];
                    uint32_t hb[4], old[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) src[u] = __ldg(p.snap_src + min(v0 + u, p.snap_n - 1));
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int v = min(v0 + u, p.snap_n - 1);
                        hb[u] = snap_one(v, src[u]);
                        old[u] = fin[(size_t)v * kLanes + lane];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (v0 + u < p.snap_n) fin[(size_t)(v0 + u) * kLanes + lane] = (HB)((old[u] & ~stop) | (hb[u] & stop));
                }
            }
            if (warp == 0) {
                for (int b = 0; b < 4 * NW; ++b)
                    if (rank == 0 && (stop >> b & 1)) iters[lane * 4 * NW + b] = (fail >> b & 1) ? -1 : it - 1;
                s_done[lane] = done | stop;
                uint32_t all_done = __reduce_and_sync(0xffffffffu, done | stop);
                if (lane == 0) s_skip = all_done & kAll;
            }
        }
        const int all = group_and(grp, ((done | stop) & kAll) == kAll);
        if (all || last) finished = true;
        else in_var = true;
        PROF_T(pt_v0);
        PROF_ADD(2, pt_c1, pt_v0);
        } else {
        PROF_T(pt_v0);
        const uint32_t vskip = s_skip;
        for (int k = 0; k < p.vc.num_classes; ++k) {
            const int deg = p.vc.deg[k], off = p.vc.off[k], cnt = p.vc.off[k + 1] - off;
            const int* vl = p.vc.var_list + off;
            const int* ve = p.vc.var_edges + p.vc.edge_off[k];
#define LDPC_VAR_CASE(D_, U_)                                                                                          \
    case D_: var_class<NW, D_, U_>(msg, hbit, inq, vl, ve, cnt, jones, D_ == 1 && d1c, vskip, gw, nw, lane, kc); break;
            #ifndef LDPC_I8_U3
#define LDPC_I8_U3 2
#endif
#ifndef LDPC_I8_U8
#define LDPC_I8_U8 1
#endif
            constexpr int U3 = NW == 1 ? 4 : LDPC_I8_U3, U8 = NW == 1 ? 2 : LDPC_I8_U8;
            switch (deg) {
                LDPC_VAR_CASE(1, U3) LDPC_VAR_CASE(2, U3) LDPC_VAR_CASE(3, U3) LDPC_VAR_CASE(4, U8)
                LDPC_VAR_CASE(5, U8) LDPC_VAR_CASE(6, U8) LDPC_VAR_CASE(7, U8) LDPC_VAR_CASE(8, U8)
                default: var_generic_class<NW>(msg, hbit, inq, g, vl, cnt, jones, gw, nw, lane, kc); break;
            }
#undef LDPC_VAR_CASE
        }
        pass_sync();
        PROF_T(pt_v1);
        PROF_ADD(3, pt_v0, pt_v1);
        PROF_ADD(4, 0, 1);
        in_var = false;
        ++it;
        }
      }
      if (kGroups == 1) {
          if (finished) break;
      } else {
          if (warp == 0 && lane == 0) s_fin[step & 1][grp] = finished ? 1 : 0;
          __syncthreads();                          // lockstep: both groups end the step together
          bool all_fin = true;
#pragma unroll
          for (int gg = 0; gg < kGroups; ++gg) all_fin = all_fin && s_fin[step & 1][gg] != 0;
          if (all_fin) break;
      }
    }
    if (C > 1) cluster.sync();          // no CTA may leave while a peer can still read its shared memory
}

template <int NW, bool AMIN, bool HLIM, int WCAP>
void launch_one(const FloodI8Launch& L, const FloodI8Params& p, cudaStream_t stream) {
    constexpr int MAXD = NW == 1 ? 10 : 8;
    constexpr int SCAP = WCAP ? WCAP : MAXD, kStages = WCAP ? 1 : 2;
    constexpr int hb = (int)sizeof(typename HBitsT<NW>::type);
#ifdef LDPC_I8_SMALL_STAGE
    constexpr size_t stage = (size_t)align128(SCAP * kLanes * NW * 4 + SCAP * kLanes * hb);
#else
    constexpr size_t stage = (size_t)align128(align128(align128(SCAP * kLanes * NW * 4 + SCAP * kLanes * hb) + kLanes * NW * 4) + kLanes * hb);
#endif
    constexpr size_t smem = (size_t)kGroups * kWarps * (kStages * stage + (size_t)kLanes * NW * 4);     // + per-warp carry slot
    // per device and cheap: set on every launch (one process may drive several GPUs)
    cudaFuncSetAttribute(flood_i8_kernel<NW, AMIN, HLIM, WCAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int C = (kGroups == 1 && L.cluster >= 1 && L.cluster <= 16) ? L.cluster : 1;
    if (C > 8) cudaFuncSetAttribute(flood_i8_kernel<NW, AMIN, HLIM, WCAP>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kCtaThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    for (;; C /= 2) {                                   // a 16-CTA cluster needs a GPC with 16 free slots: fall back to 8, 4, ..
        cfg.gridDim = dim3((unsigned)((L.num_tiles + kGroups - 1) / kGroups) * (unsigned)C);
        attr[0].val.clusterDim.x = (unsigned)C;
        int fit = 0;
        if (C == 1 || (cudaOccupancyMaxActiveClusters(&fit, flood_i8_kernel<NW, AMIN, HLIM, WCAP>, &cfg) == cudaSuccess && fit > 0)) break;
        cudaGetLastError();
    }
    cudaLaunchKernelEx(&cfg, flood_i8_kernel<NW, AMIN, HLIM, WCAP>, p);
}

template <int NW, int WCAP>
void launch_nw(const FloodI8Launch& L, const FloodI8Params& p, cudaStream_t stream) {
#ifdef LDPC_I8_BENCH_ONLY      // experiment builds (tools/build_variant.py): only the north-star instantiation
    if (NW == 4 && !L.aminstar && !L.hardlimit) launch_one<4, false, false, WCAP>(L, p, stream);
#else
    if (L.aminstar) {
        if (L.hardlimit) launch_one<NW, true, true, WCAP>(L, p, stream);
        else launch_one<NW, true, false, WCAP>(L, p, stream);
    } else {
        if (L.hardlimit) launch_one<NW, false, true, WCAP>(L, p, stream);
        else launch_one<NW, false, false, WCAP>(L, p, stream);
    }
#endif
}

}  // namespace

// one translation unit per stage capacity (flood_i8.cu: narrow, flood_i8_w16.cu, flood_i8_w32.cu): they build in parallel
#ifndef LDPC_I8_WCAP
#define LDPC_I8_WCAP 0
#endif
#if LDPC_I8_WCAP == 16
#define LDPC_I8_ENTRY launch_flood_i8_w16
#elif LDPC_I8_WCAP == 32
#define LDPC_I8_ENTRY launch_flood_i8_w32
#else
#define LDPC_I8_ENTRY launch_flood_i8_narrow
#endif

bool LDPC_I8_ENTRY(const FloodI8Launch& L, cudaStream_t stream) {
    FloodI8Params p;
    p.g = L.graph; p.vc = L.classes;
    p.msg = L.msg; p.hbit = L.hbit; p.inq = L.inq; p.raw0 = L.raw0; p.final_hard = L.final_hard; p.iters = L.iters;
    p.row_meta = L.row_meta; p.snap_src = L.snap_src; p.snap_n = L.snap_n; p.cbit = L.cbit; p.chunk_rows = L.chunk_rows; p.fuse_var_off = L.fuse_var_off; p.max_row_deg = L.graph_max_row_deg;
    p.max_iter = L.max_iter; p.num_tiles = L.num_tiles; p.jones = L.jones; p.deg1clip = L.deg1clip;
    p.c_m1 = -1; p.c_one = 1; p.c_m2 = -2; p.c_ff = 0xff; p.c_sh8 = 1 << 8; p.c_sh16 = 1 << 16; p.c_sh24 = 1 << 24;
    if (L.words_per_lane == 4) launch_nw<4, LDPC_I8_WCAP>(L, p, stream);
    else launch_nw<1, LDPC_I8_WCAP>(L, p, stream);
    LDPC_CUDA_CHECK(cudaGetLastError());
    return true;
}

#if LDPC_I8_WCAP == 0
bool launch_flood_i8_w16(const FloodI8Launch& L, cudaStream_t stream);
bool launch_flood_i8_w32(const FloodI8Launch& L, cudaStream_t stream);

// stage capacity by the widest row the register path cannot take
bool launch_flood_i8(const FloodI8Launch& L, cudaStream_t stream) {
#ifdef LDPC_I8_BENCH_ONLY
    return launch_flood_i8_narrow(L, stream);
#else
    if (L.wide_cap >= 32) return launch_flood_i8_w32(L, stream);
    if (L.wide_cap >= 16) return launch_flood_i8_w16(L, stream);
    return launch_flood_i8_narrow(L, stream);
#endif
}

int flood_i8_max_row_degree() { return kMaxGenericD; }
#endif

#if defined(LDPC_I8_PROFILE) && LDPC_I8_WCAP == 0
// experiment-only (not in include/ldpc_toolbox.h): read and reset the per-pass cycle counters
extern "C" void ldpc_toolbox_debug_i8_profile(unsigned long long* out) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_i8_prof, sizeof(unsigned long long) * 8);
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(g_i8_prof, z, sizeof(z));
}
#endif

}  // namespace ldpc
