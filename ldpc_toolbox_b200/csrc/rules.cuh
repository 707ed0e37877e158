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

// kMinstarapproxExact / kAminstarExact: the same rules with ln(1 + e^-t) evaluated by bit-exact ports of glibc's expf and
// log1pf (libm_exact.h) instead of the fast polynomial — the opt-in bit-exact mode of the f32 decoders (LDPC_B200_EXACT_LIBM=1);
// for f64 they are aliases of the plain rules.
enum RuleId { kPhi = 0, kTanh = 1, kMinstarapprox = 2, kAminstar = 3, kMinstarapproxExact = 4, kAminstarExact = 5 };
__host__ __device__ constexpr bool rule_is_minstar(int r) { return r == kMinstarapprox || r == kMinstarapproxExact; }
__host__ __device__ constexpr bool rule_is_aminstar(int r) { return r == kAminstar || r == kAminstarExact; }
__host__ __device__ constexpr bool rule_is_exact(int r) { return r == kMinstarapproxExact || r == kAminstarExact; }


#include "libm_exact.h"

template <class F> struct FMath;
template <> struct FMath<float> {
#ifdef LDPC_LIBDEVICE_TANH
    static __device__ __forceinline__ float tanh_(float x) { return tanhf(x); }
    static __device__ __forceinline__ float atanh_(float x) { return atanhf(x); }
#else
    // the f32 Tanh rule (2 atanh of a product of tanh values near +-1) gets the same bit-exact libm ports as Phi
    static __device__ __forceinline__ float tanh_(float x) { return libm_exact_tanhf(x); }
    static __device__ __forceinline__ float atanh_(float x) { return libm_exact_atanhf(x); }
#endif
    static __device__ __forceinline__ float log_(float x) { return logf(x); }
#ifdef LDPC_LIBDEVICE_PHI
    static __device__ __forceinline__ float phi_tanh_(float x) { return tanhf(x); }
    static __device__ __forceinline__ float phi_log_(float x) { return logf(x); }
#else
    static __device__ __forceinline__ float phi_tanh_(float x) { return libm_exact_tanhf(x); }      // bit-exact with the reference's libm
    static __device__ __forceinline__ float phi_log_(float x) { return libm_exact_logf(x); }
#endif
    static __device__ __forceinline__ float exp_(float x) { return expf(x); }
    static __device__ __forceinline__ float log1p_(float x) { return log1pf(x); }
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
    // the reference's own evaluation, (-t).exp().ln_1p() in f32, on bit-exact ports of the libm functions it calls
    // (a real call: inlining ~90 instructions into every step of the unrolled folds took the build from 9 to 25 minutes)
    static __device__ __noinline__ float softplus_exact(float t) { return libm_exact_log1pf(libm_exact_expf(-t)); }
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
    static __device__ __forceinline__ double softplus_exact(double t) { return log1p(exp(-t)); }
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
    } else if (rule_is_minstar(RULE)) {
        auto g = [](F a, F acc) {                              // arithmetic.rs:510
            const F sp = rule_is_exact(RULE) ? M::softplus_exact(M::abs_(a - acc)) : M::softplus_neg(M::abs_(a - acc));
            return M::max_(M::min_(a, acc) - sp, F(0));
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
            if (rule_is_exact(RULE)) return M::min_(a, b) - M::softplus_exact(M::abs_(a - b)) + M::softplus_exact(a + b);
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
