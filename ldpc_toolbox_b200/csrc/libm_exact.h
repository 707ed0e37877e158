// ldpc_toolbox_b200/csrc/libm_exact.h — bit-exact ports of the glibc 2.39 float functions the f32 Phi and Tanh rules call
// (tanhf, logf; atanhf over log1pf at the end of the file).
//
// phi(x) = -ln(tanh(x/2)) is ill-conditioned in f32 where tanh rounds towards 1: one ulp of tanhf moves phi by up to
// 6 %, so libdevice's tanhf / logf (1-2 ulp from glibc's) flipped 3 of 8192 frames at FER 3e-3.  The reference calls
// the platform libm (Rust f32::tanh / f32::ln -> tanhf / logf; reference src/decoder/arithmetic.rs:180-185); on this
// platform that is glibc 2.39, whose
//   tanhf  = fdlibm's float tanh on top of fdlibm's expm1f (pure f32 arithmetic, no tables), and
//   logf   = the table-driven double-precision algorithm of ARM's optimized routines (16-entry table, cubic),
// restated here with every operation rounded exactly as the C code does (no FMA contraction).
//
// ONE source for both sides: the GPU compiles it with the round-to-nearest intrinsics (never contracted), the host
// check (tests/libm_port_check.c, built with -ffp-contract=off by tests/test_libm_ports.py) compiles the same text
// with plain operators and compares it with the system libm on EVERY float of the domain of use
// (tanhf: all 880 803 841 floats in [2^-100, 32]; logf: all 1 115 684 864 floats in (0, 64]): zero mismatches.
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#include <math_constants.h>
#define LME_FN __device__ __forceinline__
#define LME_TAB __device__ const double
#define LME_FADD(a, b) __fadd_rn((a), (b))
#define LME_FSUB(a, b) __fsub_rn((a), (b))
#define LME_FMUL(a, b) __fmul_rn((a), (b))
#define LME_FDIV(a, b) __fdiv_rn((a), (b))
#define LME_DADD(a, b) __dadd_rn((a), (b))
#define LME_DMUL(a, b) __dmul_rn((a), (b))
#define LME_DFMA(a, b, c) __fma_rn((a), (b), (c))
#define LME_F2U(x) __float_as_uint(x)
#define LME_U2F(x) __uint_as_float(x)
#define LME_F2I_RZ(x) __float2int_rz(x)
#define LME_D2F(x) __double2float_rn(x)
#define LME_LOAD(p) __ldg(p)
#define LME_TAB64 __device__ const unsigned long long
#define LME_D2U(x) ((uint64_t)__double_as_longlong(x))
#define LME_U2D(x) __longlong_as_double((long long)(x))
#define LME_INF CUDART_INF_F
#define LME_NAN CUDART_NAN_F
#define LME_FABS(x) fabsf(x)
#else
#include <math.h>
#define LME_FN static inline
#define LME_TAB static const double
#define LME_FADD(a, b) ((float)((a) + (b)))
#define LME_FSUB(a, b) ((float)((a) - (b)))
#define LME_FMUL(a, b) ((float)((a) * (b)))
#define LME_FDIV(a, b) ((float)((a) / (b)))
#define LME_DADD(a, b) ((double)((a) + (b)))
#define LME_DMUL(a, b) ((double)((a) * (b)))
#define LME_DFMA(a, b, c) fma((a), (b), (c))
static inline uint32_t lme_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float lme_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#define LME_F2U(x) lme_f2u(x)
#define LME_U2F(x) lme_u2f(x)
#define LME_F2I_RZ(x) ((int)(x))
#define LME_D2F(x) ((float)(x))
#define LME_LOAD(p) (*(p))
#define LME_TAB64 static const unsigned long long
static inline uint64_t lme_d2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static inline double lme_u2d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
#define LME_D2U(x) lme_d2u(x)
#define LME_U2D(x) lme_u2d(x)
#define LME_INF INFINITY
#define LME_NAN NAN
#define LME_FABS(x) fabsf(x)
#endif

LME_FN float libm_exact_expm1f(float x) {
    const float one = 1.0f, huge = 1.0e+30f, tiny = 1.0e-30f, o_threshold = 8.8721679688e+01f, ln2_hi = 6.9313812256e-01f,
                ln2_lo = 9.0580006145e-06f, invln2 = 1.4426950216e+00f, Q1 = -3.3333335072e-02f, Q2 = 1.5873016091e-03f,
                Q3 = -7.9365076090e-05f, Q4 = 4.0082177293e-06f, Q5 = -2.0109921195e-07f;
    float y, hi, lo, c = 0.0f, t, e, hxs, hfx, r1;
    int k;
    uint32_t hx = LME_F2U(x);
    const uint32_t xsb = hx & 0x80000000u;
    hx &= 0x7fffffffu;
    if (hx >= 0x4195b844u) {                         // |x| >= 27 ln2
        if (hx >= 0x42b17218u) {
            if (hx > 0x7f800000u) return LME_FADD(x, x);
            if (hx == 0x7f800000u) return xsb == 0 ? x : -1.0f;
            if (x > o_threshold) return LME_FMUL(huge, huge);
        }
        if (xsb != 0) return LME_FSUB(tiny, one);
    }
    if (hx > 0x3eb17218u) {                          // |x| > 0.5 ln2
        if (hx < 0x3F851592u) {                      // |x| < 1.5 ln2
            if (xsb == 0) { hi = LME_FSUB(x, ln2_hi); lo = ln2_lo; k = 1; }
            else { hi = LME_FADD(x, ln2_hi); lo = -ln2_lo; k = -1; }
        } else {
            k = LME_F2I_RZ(LME_FADD(LME_FMUL(invln2, x), xsb == 0 ? 0.5f : -0.5f));
            t = (float)k;
            hi = LME_FSUB(x, LME_FMUL(t, ln2_hi));
            lo = LME_FMUL(t, ln2_lo);
        }
        x = LME_FSUB(hi, lo);
        c = LME_FSUB(LME_FSUB(hi, x), lo);
    } else if (hx < 0x33000000u) {                   // |x| < 2^-25
        t = LME_FADD(huge, x);
        return LME_FSUB(x, LME_FSUB(t, LME_FADD(huge, x)));
    } else {
        k = 0;
    }
    hfx = LME_FMUL(0.5f, x);
    hxs = LME_FMUL(x, hfx);
    r1 = LME_FADD(one, LME_FMUL(hxs, LME_FADD(Q1, LME_FMUL(hxs, LME_FADD(Q2, LME_FMUL(hxs, LME_FADD(Q3, LME_FMUL(hxs, LME_FADD(Q4, LME_FMUL(hxs, Q5))))))))));
    t = LME_FSUB(3.0f, LME_FMUL(r1, hfx));
    e = LME_FMUL(hxs, LME_FDIV(LME_FSUB(r1, t), LME_FSUB(6.0f, LME_FMUL(x, t))));
    if (k == 0) return LME_FSUB(x, LME_FSUB(LME_FMUL(x, e), hxs));
    e = LME_FSUB(LME_FMUL(x, LME_FSUB(e, c)), c);
    e = LME_FSUB(e, hxs);
    if (k == -1) return LME_FSUB(LME_FMUL(0.5f, LME_FSUB(x, e)), 0.5f);
    if (k == 1) {
        if (x < -0.25f) return LME_FMUL(-2.0f, LME_FSUB(e, LME_FADD(x, 0.5f)));
        return LME_FADD(one, LME_FMUL(2.0f, LME_FSUB(x, e)));
    }
    if (k <= -2 || k > 56) {
        y = LME_FSUB(one, LME_FSUB(e, x));
        y = LME_U2F(LME_F2U(y) + ((uint32_t)k << 23));
        return LME_FSUB(y, one);
    }
    if (k < 23) {
        t = LME_U2F(0x3f800000u - (0x1000000u >> k));
        y = LME_FSUB(t, LME_FSUB(e, x));
        y = LME_U2F(LME_F2U(y) + ((uint32_t)k << 23));
    } else {
        t = LME_U2F((uint32_t)(0x7f - k) << 23);
        y = LME_FSUB(x, LME_FADD(e, t));
        y = LME_FADD(y, one);
        y = LME_U2F(LME_F2U(y) + ((uint32_t)k << 23));
    }
    return y;
}

LME_FN float libm_exact_tanhf(float x) {
    const float one = 1.0f, tiny = 1.0e-30f;
    const uint32_t jx = LME_F2U(x), ix = jx & 0x7fffffffu;
    float t, z;
    if (!(ix < 0x7f800000u)) return (int32_t)jx >= 0 ? LME_FADD(LME_FDIV(one, x), one) : LME_FSUB(LME_FDIV(one, x), one);
    if (ix < 0x41b00000u) {                          // |x| < 22
        if (ix == 0) return x;
        if (ix < 0x24000000u) return LME_FMUL(x, LME_FADD(one, x));
        if (ix >= 0x3f800000u) {                     // |x| >= 1
            t = libm_exact_expm1f(LME_FMUL(2.0f, LME_FABS(x)));
            z = LME_FSUB(one, LME_FDIV(2.0f, LME_FADD(t, 2.0f)));
        } else {
            t = libm_exact_expm1f(LME_FMUL(-2.0f, LME_FABS(x)));
            z = LME_FDIV(-t, LME_FADD(t, 2.0f));
        }
    } else {
        z = LME_FSUB(one, tiny);
    }
    return (int32_t)jx >= 0 ? z : -z;
}

// {1/c, ln c} for the 16 sub-intervals of [0.7, 1.4) and the cubic of ln(1 + r) — glibc's __logf_data
LME_TAB kLibmExactLogfTab[32] = {
    0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2, 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2, 0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2,
    0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3, 0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3, 0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3,
    0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4, 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4, 0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5,
    0x1p+0, 0x0p+0, 0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5, 0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4,
    0x1.b2036576afce6p-1, 0x1.526e57720db08p-3, 0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3, 0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2,
    0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2};

LME_FN float libm_exact_logf(float x) {
    const double A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2, Ln2 = 0x1.62e42fefa39efp-1;
    uint32_t ix = LME_F2U(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2 == 0) return -LME_INF;
        if (ix == 0x7f800000u) return x;
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return LME_NAN;
        ix = LME_F2U(LME_FMUL(x, 0x1p23f));  // subnormal: normalise
        ix -= 23u << 23;
    }
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15u), k = (int32_t)tmp >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    const double invc = LME_LOAD(&kLibmExactLogfTab[2 * i]), logc = LME_LOAD(&kLibmExactLogfTab[2 * i + 1]), z = (double)LME_U2F(iz);
    const double r = LME_DADD(LME_DMUL(z, invc), -1.0), y0 = LME_DADD(logc, LME_DMUL((double)k, Ln2)), r2 = LME_DMUL(r, r);
    double y = LME_DADD(LME_DMUL(A1, r), A2);
    y = LME_DADD(LME_DMUL(A0, r2), y);
    y = LME_DADD(LME_DMUL(y, r2), LME_DADD(y0, r));
    return LME_D2F(y);
}

// ---- atanhf (glibc 2.39: sysdeps/ieee754/flt-32/e_atanhf.c on top of fdlibm's s_log1pf.c), used by the f32 Tanh rule
// (reference src/decoder/arithmetic.rs:347-379: 2 atanh(prod of tanh)).  Same conventions as above.
LME_FN float libm_exact_log1pf(float x) {
    const float ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f, two25 = 3.355443200e+07f, Lp1 = 6.6666668653e-01f,
                Lp2 = 4.0000000596e-01f, Lp3 = 2.8571429849e-01f, Lp4 = 2.2222198546e-01f, Lp5 = 1.8183572590e-01f,
                Lp6 = 1.5313838422e-01f, Lp7 = 1.4798198640e-01f, zero = 0.0f;
    float hfsq, f = 0.0f, c = 0.0f, s, z, R, u;
    int32_t k, hx, hu = 0, ax;
    hx = (int32_t)LME_F2U(x);
    ax = hx & 0x7fffffff;
    k = 1;
    if (hx < 0x3ed413d7) {                           // x < 0.41422
        if (ax >= 0x3f800000) {                      // x <= -1.0
            if (x == -1.0f) return LME_FDIV(-two25, zero);
            return LME_FDIV(LME_FSUB(x, x), LME_FSUB(x, x));
        }
        if (ax < 0x31000000) {                       // |x| < 2^-29
            if (ax < 0x24800000) return x;           // |x| < 2^-54
            return LME_FSUB(x, LME_FMUL(LME_FMUL(x, x), 0.5f));
        }
        if (hx > 0 || hx <= (int32_t)0xbe95f61f) { k = 0; f = x; hu = 1; }      // -0.2929 < x < 0.41422
    }
    if (hx >= 0x7f800000) return LME_FADD(x, x);
    if (k != 0) {
        if (hx < 0x5a000000) {
            u = LME_FADD(1.0f, x);
            hu = (int32_t)LME_F2U(u);
            k = (hu >> 23) - 127;
            c = (k > 0) ? LME_FSUB(1.0f, LME_FSUB(u, x)) : LME_FSUB(x, LME_FSUB(u, 1.0f));      // correction term
            c = LME_FDIV(c, u);
        } else {
            u = x;
            hu = (int32_t)LME_F2U(u);
            k = (hu >> 23) - 127;
            c = 0.0f;
        }
        hu &= 0x007fffff;
        if (hu < 0x3504f7) {
            u = LME_U2F((uint32_t)hu | 0x3f800000u);          // normalize u
        } else {
            k += 1;
            u = LME_U2F((uint32_t)hu | 0x3f000000u);          // normalize u/2
            hu = (0x00800000 - hu) >> 2;
        }
        f = LME_FSUB(u, 1.0f);
    }
    hfsq = LME_FMUL(LME_FMUL(0.5f, f), f);
    if (hu == 0) {                                   // |f| < 2^-20
        if (f == zero) {
            if (k == 0) return zero;
            c = LME_FADD(c, LME_FMUL((float)k, ln2_lo));
            return LME_FADD(LME_FMUL((float)k, ln2_hi), c);
        }
        R = LME_FMUL(hfsq, LME_FSUB(1.0f, LME_FMUL(0.66666666666666666f, f)));
        if (k == 0) return LME_FSUB(f, R);
        return LME_FSUB(LME_FMUL((float)k, ln2_hi), LME_FSUB(LME_FSUB(R, LME_FADD(LME_FMUL((float)k, ln2_lo), c)), f));
    }
    s = LME_FDIV(f, LME_FADD(2.0f, f));
    z = LME_FMUL(s, s);
    R = LME_FMUL(z, LME_FADD(Lp1, LME_FMUL(z, LME_FADD(Lp2, LME_FMUL(z, LME_FADD(Lp3, LME_FMUL(z, LME_FADD(Lp4, LME_FMUL(z, LME_FADD(Lp5, LME_FMUL(z, LME_FADD(Lp6, LME_FMUL(z, Lp7)))))))))))));
    if (k == 0) return LME_FSUB(f, LME_FSUB(hfsq, LME_FMUL(s, LME_FADD(hfsq, R))));
    return LME_FSUB(LME_FMUL((float)k, ln2_hi),
                    LME_FSUB(LME_FSUB(hfsq, LME_FADD(LME_FMUL(s, LME_FADD(hfsq, R)), LME_FADD(LME_FMUL((float)k, ln2_lo), c))), f));
}

LME_FN float libm_exact_atanhf(float x) {
    const float xa = LME_FABS(x);
    float t;
    if (xa < 0.5f) {
        if (xa < 0x1.0p-28f) return x;
        t = LME_FADD(xa, xa);
        t = LME_FMUL(0.5f, libm_exact_log1pf(LME_FADD(t, LME_FDIV(LME_FMUL(t, xa), LME_FSUB(1.0f, xa)))));
    } else if (xa < 1.0f) {
        t = LME_FMUL(0.5f, libm_exact_log1pf(LME_FDIV(LME_FADD(xa, xa), LME_FSUB(1.0f, xa))));
    } else {
        if (xa > 1.0f) return LME_FDIV(LME_FSUB(x, x), LME_FSUB(x, x));
        return LME_FDIV(x, 0.0f);
    }
    return LME_U2F((LME_F2U(t) & 0x7fffffffu) | (LME_F2U(x) & 0x80000000u));
}

// ---- expf (glibc 2.39: sysdeps/ieee754/flt-32/e_expf.c, ARM optimized routines: 2^(k/32) table, cubic in double),
// with log1pf above the exact ln(1 + e^-t) of the f32 Min*-approx / A-Min* rules (reference src/decoder/arithmetic.rs:510,
// :965-966) in their opt-in bit-exact mode.  kLibmExactExp2fTab[i] = bits(2^(i/32)) - (i << 47).
LME_TAB64 kLibmExactExp2fTab[32] = {
0x3ff0000000000000ull,
0x3fefd9b0d3158574ull,
0x3fefb5586cf9890full,
0x3fef9301d0125b51ull,
0x3fef72b83c7d517bull,
0x3fef54873168b9aaull,
0x3fef387a6e756238ull,
0x3fef1e9df51fdee1ull,
0x3fef06fe0a31b715ull,
0x3feef1a7373aa9cbull,
0x3feedea64c123422ull,
0x3feece086061892dull,
0x3feebfdad5362a27ull,
0x3feeb42b569d4f82ull,
0x3feeab07dd485429ull,
0x3feea47eb03a5585ull,
0x3feea09e667f3bcdull,
0x3fee9f75e8ec5f74ull,
0x3feea11473eb0187ull,
0x3feea589994cce13ull,
0x3feeace5422aa0dbull,
0x3feeb737b0cdc5e5ull,
0x3feec49182a3f090ull,
0x3feed503b23e255dull,
0x3feee89f995ad3adull,
0x3feeff76f2fb5e47ull,
0x3fef199bdd85529cull,
0x3fef3720dcef9069ull,
0x3fef5818dcfba487ull,
0x3fef7c97337b9b5full,
0x3fefa4afa2a490daull,
0x3fefd0765b6e4540ull};

LME_FN float libm_exact_expf(float x) {
    const double InvLn2N = 0x1.71547652b82fep+5, Shift = 0x1.8p+52, C0 = 0x1.c6af84b912394p-20, C1 = 0x1.ebfce50fac4f3p-13,
                 C2 = 0x1.62e42ff0c52d6p-6;
    const uint32_t ux = LME_F2U(x), abstop = (ux >> 20) & 0x7ff;
    if (abstop >= (0x42b00000u >> 20)) {             // |x| >= 88 or NaN
        if (ux == 0xff800000u) return 0.0f;
        if (abstop >= (0x7f800000u >> 20)) return LME_FADD(x, x);
        if (x > 0x1.62e42ep6f) return LME_INF;       // overflow
        if (x < -0x1.9fe368p6f) return 0.0f;         // underflow
    }
    const double xd = (double)x;
    double z = LME_DMUL(InvLn2N, xd);
    double kd = LME_DADD(z, Shift);
    const uint64_t ki = LME_D2U(kd);
    kd = LME_DADD(kd, -Shift);
    // x86-64 glibc dispatches expf to its FMA build on every CPU that has FMA (sysdeps/x86_64/fpu/multiarch/e_expf.c); there the
    // compiler contracts r = InvLn2N x - kd and the three polynomial steps into fused multiply-adds, and two floats out of
    // 2.2 billion round differently without the first of them (x = 0x1.04845ep+5 and -0x1.f8cbb2p+5)
    const double r = LME_DFMA(InvLn2N, xd, -kd);
    uint64_t t = LME_LOAD(&kLibmExactExp2fTab[ki & 31]);
    t += ki << 47;
    const double s = LME_U2D(t);
    z = LME_DFMA(C0, r, C1);
    const double r2 = LME_DMUL(r, r);
    double y = LME_DFMA(C2, r, 1.0);
    y = LME_DFMA(z, r2, y);
    y = LME_DMUL(y, s);
    return LME_D2F(y);
}
