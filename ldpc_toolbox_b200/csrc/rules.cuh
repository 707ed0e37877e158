// ldpc_toolbox_b200/csrc/rules.cuh — per-frame check-node rules shared by the generic flooding
// kernel (K2) and the horizontal-layered kernel (K3).
//
// Each rule turns the d incoming variable->check values of ONE frame (x[0..d), row order) into the d
// outgoing check->variable values, exactly following the evaluation order of the reference:
//   Phi            reference src/decoder/arithmetic.rs:214-246
//   Tanh           reference src/decoder/arithmetic.rs:347-379
//   Minstarapprox  reference src/decoder/arithmetic.rs:487-521 (float), :718-754 (i8)
//   Aminstar       reference src/decoder/arithmetic.rs:942-999 (float), :1130-1192 (i8)
// Float transcendental results come from CUDA's libdevice instead of the host libm, so float
// decoders are tolerance-parity; the i8 rules are bit-exact.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <cstdint>

namespace ldpc {

constexpr int kRuleMaxD = 32;        // largest check degree handled by the generic kernels

enum RuleId { kPhi = 0, kTanh = 1, kMinstarapprox = 2, kAminstar = 3 };


// ---- bit-exact ports of the two glibc 2.39 float functions the f32 Phi rule calls ---------------------------------
// phi(x) = -ln(tanh(x/2)) is ill-conditioned in f32 where tanh rounds towards 1: one ulp of tanhf moves phi by up to
// 6 %, so libdevice's tanhf / logf (1-2 ulp from glibc's) flipped 3 of 8192 frames at FER 3e-3.  The reference calls
// the platform libm (Rust f32::tanh / f32::ln -> tanhf / logf); on this platform that is glibc 2.39, whose
//   tanhf  = fdlibm's float tanh on top of fdlibm's expm1f (pure f32 arithmetic, no tables), and
//   logf   = the table-driven double-precision algorithm of ARM's optimized routines (16-entry table, cubic),
// restated here with every operation rounded exactly as the C code does (no FMA contraction: __f*_rn / __d*_rn).
// Both were compared with the system libm on EVERY float of their domain of use before going to the GPU
// (tanhf: all 880 803 841 floats in [2^-100, 32]; logf: all 1 115 684 864 floats in (0, 64]): zero mismatches.
__device__ __forceinline__ float glibc_expm1f(float x) {
    const float one = 1.0f, huge = 1.0e+30f, tiny = 1.0e-30f, o_threshold = 8.8721679688e+01f, ln2_hi = 6.9313812256e-01f,
                ln2_lo = 9.0580006145e-06f, invln2 = 1.4426950216e+00f, Q1 = -3.3333335072e-02f, Q2 = 1.5873016091e-03f,
                Q3 = -7.9365076090e-05f, Q4 = 4.0082177293e-06f, Q5 = -2.0109921195e-07f;
    float y, hi, lo, c = 0.0f, t, e, hxs, hfx, r1;
    int k;
    uint32_t hx = __float_as_uint(x);
    const uint32_t xsb = hx & 0x80000000u;
    hx &= 0x7fffffffu;
    if (hx >= 0x4195b844u) {                         // |x| >= 27 ln2
        if (hx >= 0x42b17218u) {
            if (hx > 0x7f800000u) return __fadd_rn(x, x);
            if (hx == 0x7f800000u) return xsb == 0 ? x : -1.0f;
            if (x > o_threshold) return __fmul_rn(huge, huge);
        }
        if (xsb != 0) return __fsub_rn(tiny, one);
    }
    if (hx > 0x3eb17218u) {                          // |x| > 0.5 ln2
        if (hx < 0x3F851592u) {                      // |x| < 1.5 ln2
            if (xsb == 0) { hi = __fsub_rn(x, ln2_hi); lo = ln2_lo; k = 1; }
            else { hi = __fadd_rn(x, ln2_hi); lo = -ln2_lo; k = -1; }
        } else {
            k = __float2int_rz(__fadd_rn(__fmul_rn(invln2, x), xsb == 0 ? 0.5f : -0.5f));
            t = (float)k;
            hi = __fsub_rn(x, __fmul_rn(t, ln2_hi));
            lo = __fmul_rn(t, ln2_lo);
        }
        x = __fsub_rn(hi, lo);
        c = __fsub_rn(__fsub_rn(hi, x), lo);
    } else if (hx < 0x33000000u) {                   // |x| < 2^-25
        t = __fadd_rn(huge, x);
        return __fsub_rn(x, __fsub_rn(t, __fadd_rn(huge, x)));
    } else {
        k = 0;
    }
    hfx = __fmul_rn(0.5f, x);
    hxs = __fmul_rn(x, hfx);
    r1 = __fadd_rn(one, __fmul_rn(hxs, __fadd_rn(Q1, __fmul_rn(hxs, __fadd_rn(Q2, __fmul_rn(hxs, __fadd_rn(Q3, __fmul_rn(hxs, __fadd_rn(Q4, __fmul_rn(hxs, Q5))))))))));
    t = __fsub_rn(3.0f, __fmul_rn(r1, hfx));
    e = __fmul_rn(hxs, __fdiv_rn(__fsub_rn(r1, t), __fsub_rn(6.0f, __fmul_rn(x, t))));
    if (k == 0) return __fsub_rn(x, __fsub_rn(__fmul_rn(x, e), hxs));
    e = __fsub_rn(__fmul_rn(x, __fsub_rn(e, c)), c);
    e = __fsub_rn(e, hxs);
    if (k == -1) return __fsub_rn(__fmul_rn(0.5f, __fsub_rn(x, e)), 0.5f);
    if (k == 1) {
        if (x < -0.25f) return __fmul_rn(-2.0f, __fsub_rn(e, __fadd_rn(x, 0.5f)));
        return __fadd_rn(one, __fmul_rn(2.0f, __fsub_rn(x, e)));
    }
    if (k <= -2 || k > 56) {
        y = __fsub_rn(one, __fsub_rn(e, x));
        y = __uint_as_float(__float_as_uint(y) + ((uint32_t)k << 23));
        return __fsub_rn(y, one);
    }
    if (k < 23) {
        t = __uint_as_float(0x3f800000u - (0x1000000u >> k));
        y = __fsub_rn(t, __fsub_rn(e, x));
        y = __uint_as_float(__float_as_uint(y) + ((uint32_t)k << 23));
    } else {
        t = __uint_as_float((uint32_t)(0x7f - k) << 23);
        y = __fsub_rn(x, __fadd_rn(e, t));
        y = __fadd_rn(y, one);
        y = __uint_as_float(__float_as_uint(y) + ((uint32_t)k << 23));
    }
    return y;
}

__device__ __forceinline__ float glibc_tanhf(float x) {
    const float one = 1.0f, tiny = 1.0e-30f;
    const uint32_t jx = __float_as_uint(x), ix = jx & 0x7fffffffu;
    float t, z;
    if (!(ix < 0x7f800000u)) return (int32_t)jx >= 0 ? __fadd_rn(__fdiv_rn(one, x), one) : __fsub_rn(__fdiv_rn(one, x), one);
    if (ix < 0x41b00000u) {                          // |x| < 22
        if (ix == 0) return x;
        if (ix < 0x24000000u) return __fmul_rn(x, __fadd_rn(one, x));
        if (ix >= 0x3f800000u) {                     // |x| >= 1
            t = glibc_expm1f(__fmul_rn(2.0f, fabsf(x)));
            z = __fsub_rn(one, __fdiv_rn(2.0f, __fadd_rn(t, 2.0f)));
        } else {
            t = glibc_expm1f(__fmul_rn(-2.0f, fabsf(x)));
            z = __fdiv_rn(-t, __fadd_rn(t, 2.0f));
        }
    } else {
        z = __fsub_rn(one, tiny);
    }
    return (int32_t)jx >= 0 ? z : -z;
}

// {1/c, ln c} for the 16 sub-intervals of [0.7, 1.4) and the cubic of ln(1 + r) — glibc's __logf_data
__device__ const double kGlibcLogfTab[32] = {
    0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2, 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2, 0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2,
    0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3, 0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3, 0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3,
    0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4, 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4, 0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5,
    0x1p+0, 0x0p+0, 0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5, 0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4,
    0x1.b2036576afce6p-1, 0x1.526e57720db08p-3, 0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3, 0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2,
    0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2};

__device__ __forceinline__ float glibc_logf(float x) {
    const double A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2, Ln2 = 0x1.62e42fefa39efp-1;
    uint32_t ix = __float_as_uint(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2 == 0) return -CUDART_INF_F;
        if (ix == 0x7f800000u) return x;
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return CUDART_NAN_F;
        ix = __float_as_uint(__fmul_rn(x, 0x1p23f));  // subnormal: normalise
        ix -= 23u << 23;
    }
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15u), k = (int32_t)tmp >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    const double invc = __ldg(&kGlibcLogfTab[2 * i]), logc = __ldg(&kGlibcLogfTab[2 * i + 1]), z = (double)__uint_as_float(iz);
    const double r = __dadd_rn(__dmul_rn(z, invc), -1.0), y0 = __dadd_rn(logc, __dmul_rn((double)k, Ln2)), r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(A1, r), A2);
    y = __dadd_rn(__dmul_rn(A0, r2), y);
    y = __dadd_rn(__dmul_rn(y, r2), __dadd_rn(y0, r));
    return __double2float_rn(y);
}

template <class F> struct FMath;
template <> struct FMath<float> {
    static __device__ __forceinline__ float tanh_(float x) { return tanhf(x); }
    static __device__ __forceinline__ float log_(float x) { return logf(x); }
#ifdef LDPC_LIBDEVICE_PHI
    static __device__ __forceinline__ float phi_tanh_(float x) { return tanhf(x); }
    static __device__ __forceinline__ float phi_log_(float x) { return logf(x); }
#else
    static __device__ __forceinline__ float phi_tanh_(float x) { return glibc_tanhf(x); }      // bit-exact with the reference's libm
    static __device__ __forceinline__ float phi_log_(float x) { return glibc_logf(x); }
#endif
    static __device__ __forceinline__ float exp_(float x) { return expf(x); }
    static __device__ __forceinline__ float log1p_(float x) { return log1pf(x); }
    static __device__ __forceinline__ float atanh_(float x) { return atanhf(x); }
    static __device__ __forceinline__ float abs_(float x) { return fabsf(x); }
    static __device__ __forceinline__ float max_(float a, float b) { return fmaxf(a, b); }
    static __device__ __forceinline__ float min_(float a, float b) { return fminf(a, b); }
    static __device__ __forceinline__ float tanh_clamp() { return 9.0f; }     // arithmetic.rs:435
    // ln(1 + e^-t), t >= 0 — the correction term of min* (arithmetic.rs:510, :965).  libdevice's log1pf(expf(-t)) costs
    // ~40 instructions and dominated the f32 min* kernels; this is one MUFU.EX2 and ten FMAs: e = 2^(-t log2 e), then
    // ln(1 + e) = e p(e) with a degree-9 minimax polynomial on [0, 1].  Absolute error <= 2.5e-7 (the libm pair the
    // reference calls is within ~1e-7 of the true value), far below the f32 resolution of the messages it is
    // subtracted from; -DLDPC_EXACT_SOFTPLUS restores the libdevice pair.  Parity at scale: tests/test_gpu_parity_scale.py.
    static __device__ __forceinline__ float softplus_neg(float t) {
#ifdef LDPC_EXACT_SOFTPLUS
        return log1pf(expf(-t));
#else
        const float e = exp2f(-1.4426950408889634f * t);
        float p = -0.003256378462538123f;
        p = fmaf(p, e, 0.019907161593437195f);
        p = fmaf(p, e, -0.057064201682806015f);
        p = fmaf(p, e, 0.10614264756441116f);
        p = fmaf(p, e, -0.15311862528324127f);
        p = fmaf(p, e, 0.19678117334842682f);
        p = fmaf(p, e, -0.24954558908939362f);
        p = fmaf(p, e, 0.33330005407333374f);
        p = fmaf(p, e, -0.4999990463256836f);
        p = fmaf(p, e, 1.0f);
        return p * e;
#endif
    }
};
template <> struct FMath<double> {
    static __device__ __forceinline__ double tanh_(double x) { return tanh(x); }
    static __device__ __forceinline__ double log_(double x) { return log(x); }
    static __device__ __forceinline__ double phi_tanh_(double x) { return tanh(x); }
    static __device__ __forceinline__ double phi_log_(double x) { return log(x); }
    static __device__ __forceinline__ double exp_(double x) { return exp(x); }
    static __device__ __forceinline__ double log1p_(double x) { return log1p(x); }
    static __device__ __forceinline__ double atanh_(double x) { return atanh(x); }
    static __device__ __forceinline__ double softplus_neg(double t) { return log1p(exp(-t)); }
    static __device__ __forceinline__ double abs_(double x) { return fabs(x); }
    static __device__ __forceinline__ double max_(double a, double b) { return fmax(a, b); }
    static __device__ __forceinline__ double min_(double a, double b) { return fmin(a, b); }
    static __device__ __forceinline__ double tanh_clamp() { return 18.0; }    // arithmetic.rs:433
};

// ---- float rules: x in, out out (may not alias), scratch has room for d values -----------------
// DT > 0: the degree is the compile-time constant DT (every loop unrolls and x / out / scratch stay
// in registers); DT == 0: run-time degree d_rt.
template <class F, int RULE, int DT = 0>
__device__ __forceinline__ void check_rule_float(const F* x, int d_rt, F* out, F* scratch) {
    using M = FMath<F>;
    const int d = DT > 0 ? DT : d_rt;
    if (RULE == kPhi) {
        auto phi = [](F v) {                                   // arithmetic.rs:180-185
            v = M::max_(v, F(1e-30));
            return -M::phi_log_(M::phi_tanh_(F(0.5) * v));
        };
        unsigned sign = 0;
        F sum = F(0);
        _Pragma("unroll") for (int i = 0; i < d; ++i) {
            F p = phi(M::abs_(x[i]));
            scratch[i] = p;
            sum += p;
            if (x[i] < F(0)) sign ^= 1u;
        }
        _Pragma("unroll") for (int i = 0; i < d; ++i) {
            F y = phi(sum - scratch[i]);
            unsigned s = x[i] < F(0) ? (sign ^ 1u) : sign;
            out[i] = s == 0 ? y : -y;
        }
    } else if (RULE == kTanh) {
        const F c = M::tanh_clamp();
        _Pragma("unroll") for (int i = 0; i < d; ++i) {
            F h = F(0.5) * x[i];
            h = h < -c ? -c : (h > c ? c : h);                 // Rust clamp (NaN propagates)
            scratch[i] = M::tanh_(h);
        }
        _Pragma("unroll") for (int j = 0; j < d; ++j) {
            F prod = F(1);
            _Pragma("unroll") for (int i = 0; i < d; ++i)
                if (i != j) prod *= scratch[i];
            out[j] = F(2) * M::atanh_(prod);
        }
    } else if (RULE == kMinstarapprox) {
        auto g = [](F a, F acc) {                              // arithmetic.rs:510
            return M::max_(M::min_(a, acc) - M::softplus_neg(M::abs_(a - acc)), F(0));
        };
        // shared prefix P_j = fold(|x_0| .. |x_{j-1}|); the remaining terms are folded per output
        F P = F(0);
        _Pragma("unroll") for (int j = 0; j < d; ++j) {
            unsigned sign = 0;
            _Pragma("unroll") for (int i = 0; i < d; ++i)
                if (i != j && x[i] < F(0)) sign ^= 1u;
            F acc = P;
            bool have = j > 0;
            _Pragma("unroll") for (int i = j + 1; i < d; ++i) {
                F a = M::abs_(x[i]);
                acc = have ? g(a, acc) : a;
                have = true;
            }
            out[j] = sign == 0 ? acc : -acc;
            F aj = M::abs_(x[j]);
            P = j == 0 ? aj : g(aj, P);
        }
    } else {                                                   // A-Min*
        auto h = [](F a, F b) {                                // arithmetic.rs:965-966
            return M::min_(a, b) - M::softplus_neg(M::abs_(a - b)) + M::softplus_neg(a + b);
        };
        int arg = 0;
        F best = M::abs_(x[0]);
        _Pragma("unroll") for (int i = 1; i < d; ++i) {
            F a = M::abs_(x[i]);
            if (a < best) { best = a; arg = i; }               // first minimum
        }
        unsigned sign = 0;
        bool have = false;
        F delta = F(0);
        _Pragma("unroll") for (int j = 0; j < d; ++j) {
            if (x[j] < F(0)) sign ^= 1u;
            if (j != arg) {
                F a = M::abs_(x[j]);
                delta = have ? h(a, delta) : a;
                have = true;
            }
        }
        F d2 = h(delta, best);                                 // best == |x[arg]|
        _Pragma("unroll") for (int j = 0; j < d; ++j) {
            F mag = j == arg ? delta : d2;
            bool neg = (sign != 0) ^ (x[j] < F(0));
            out[j] = neg ? -mag : mag;
        }
    }
}

// ---- int8 rules (used by the layered kernel; the flooding i8 kernel has its own packed path) ---
struct I8Tables {
    int8_t U[256];      // U[d + 127] = min(d, 0) - T[|d|]
    int8_t Tp[128];     // T[t] = round(8 ln(1 + e^{-t/8}))
};

__device__ __forceinline__ int i8_table_T(int t) {
    return (t < 1) + (t < 3) + (t < 5) + (t < 9) + (t < 13) + (t < 22);
}

__device__ __forceinline__ void i8_tables_init(I8Tables& tb) {
    for (int i = threadIdx.x; i < 255; i += blockDim.x) {
        int d = i - 127;
        tb.U[i] = (int8_t)(min(d, 0) - i8_table_T(abs(d)));
    }
    for (int i = threadIdx.x; i < 128; i += blockDim.x) tb.Tp[i] = (int8_t)i8_table_T(i);
}

__device__ __forceinline__ int i8_clip(int x) { return x >= 127 ? 127 : (x <= -127 ? -127 : x); }   // arithmetic.rs:609-617

template <int RULE, bool HLIM, int DT = 0>
__device__ __forceinline__ void check_rule_i8(const int* x, int d_rt, int* out, const I8Tables& tb) {
    const int d = DT > 0 ? DT : d_rt;
    auto hl = [](int m) { return HLIM ? (m >= 100 ? 127 : m) : m; };
    if (RULE == kMinstarapprox) {
        auto g = [&](int a, int acc) { return max(acc + (int)tb.U[a - acc + 127], 0); };
        int P = 0;
        _Pragma("unroll") for (int j = 0; j < d; ++j) {
            unsigned sign = 0;
            _Pragma("unroll") for (int i = 0; i < d; ++i)
                if (i != j && x[i] < 0) sign ^= 1u;
            int acc = P;
            bool have = j > 0;
            _Pragma("unroll") for (int i = j + 1; i < d; ++i) {
                int a = abs(x[i]);
                acc = have ? g(a, acc) : a;
                have = true;
            }
            int m = hl(acc);
            out[j] = sign == 0 ? m : -m;
            int aj = abs(x[j]);
            P = j == 0 ? aj : g(aj, P);
        }
    } else {
        auto h = [&](int a, int b) { return max(b + (int)tb.U[a - b + 127] + (int)tb.Tp[min(a + b, 127)], 0); };
        int arg = 0, best = abs(x[0]);
        _Pragma("unroll") for (int i = 1; i < d; ++i) {
            int a = abs(x[i]);
            if (a < best) { best = a; arg = i; }
        }
        unsigned sign = 0;
        bool have = false;
        int delta = 0;
        _Pragma("unroll") for (int j = 0; j < d; ++j) {
            if (x[j] < 0) sign ^= 1u;
            if (j != arg) {
                int a = abs(x[j]);
                delta = have ? h(a, delta) : a;
                have = true;
            }
        }
        int d2 = hl(h(delta, best));                           // best == |x[arg]|
        int d1 = hl(delta);
        _Pragma("unroll") for (int j = 0; j < d; ++j) {
            int mag = j == arg ? d1 : d2;
            bool neg = (sign != 0) ^ (x[j] < 0);
            out[j] = neg ? -mag : mag;
        }
    }
}

}  // namespace ldpc
