/* tests/libm_port_check.c — host check of ldpc_toolbox_b200/csrc/libm_exact.h (test infrastructure): the very text the
 * GPU compiles, built here with plain operators and -ffp-contract=off, against the system libm on every float of the
 * domain the f32 Phi rule uses.  Prints the mismatch counts; tests/test_libm_ports.py asserts they are zero.
 *   gcc -O2 -ffp-contract=off -fopenmp tests/libm_port_check.c -o check -lm && ./check [stride] */
#include <stdio.h>
#include <stdlib.h>
#include "../ldpc_toolbox_b200/csrc/libm_exact.h"

int main(int argc, char** argv) {
    const uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1u;
    long bad_tanh = 0, bad_log = 0, n_tanh = 0, n_log = 0;
    const uint32_t t_lo = lme_f2u(0x1p-100f), t_hi = lme_f2u(32.0f), l_hi = lme_f2u(64.0f);
#pragma omp parallel for reduction(+ : bad_tanh, n_tanh) schedule(static)
    for (uint32_t u = t_lo; u <= t_hi; u += stride) {
        const float x = lme_u2f(u);
        bad_tanh += lme_f2u(tanhf(x)) != lme_f2u(libm_exact_tanhf(x));
        bad_tanh += lme_f2u(tanhf(-x)) != lme_f2u(libm_exact_tanhf(-x));
        ++n_tanh;
    }
#pragma omp parallel for reduction(+ : bad_log, n_log) schedule(static)
    for (uint32_t u = 1; u <= l_hi; u += stride) {
        const float x = lme_u2f(u);
        bad_log += lme_f2u(logf(x)) != lme_f2u(libm_exact_logf(x));
        ++n_log;
    }
    long bad_atanh = 0, n_atanh = 0;
    const uint32_t a_hi = lme_f2u(1.0f);
#pragma omp parallel for reduction(+ : bad_atanh, n_atanh) schedule(static)
    for (uint32_t u = 0; u <= a_hi; u += stride) {
        const float x = lme_u2f(u);
        bad_atanh += lme_f2u(atanhf(x)) != lme_f2u(libm_exact_atanhf(x));
        bad_atanh += lme_f2u(atanhf(-x)) != lme_f2u(libm_exact_atanhf(-x));
        ++n_atanh;
    }
    long bad_exp = 0, n_exp = 0, bad_l1p = 0, n_l1p = 0;
    const uint32_t e_hi = lme_f2u(104.0f), p_hi = lme_f2u(2.0f);
#pragma omp parallel for reduction(+ : bad_exp, n_exp) schedule(static)
    for (uint32_t u = 0; u <= e_hi; u += stride) {                 /* exp(-t), t in [0, 104]; and exp(+t) up to overflow */
        const float x = lme_u2f(u);
        bad_exp += lme_f2u(expf(-x)) != lme_f2u(libm_exact_expf(-x));
        bad_exp += lme_f2u(expf(x)) != lme_f2u(libm_exact_expf(x));
        ++n_exp;
    }
#pragma omp parallel for reduction(+ : bad_l1p, n_l1p) schedule(static)
    for (uint32_t u = 0; u <= p_hi; u += stride) {                 /* log1p(e), e = exp(-t) in [0, 1] (checked up to 2) */
        const float x = lme_u2f(u);
        bad_l1p += lme_f2u(log1pf(x)) != lme_f2u(libm_exact_log1pf(x));
        ++n_l1p;
    }
    printf("{\"expf_checked\": %ld, \"expf_mismatches\": %ld, \"log1pf_checked\": %ld, \"log1pf_mismatches\": %ld}\n", n_exp, bad_exp, n_l1p, bad_l1p);
    printf("{\"atanhf_checked\": %ld, \"atanhf_mismatches\": %ld}\n", n_atanh, bad_atanh);
    printf("{\"tanhf_checked\": %ld, \"tanhf_mismatches\": %ld, \"logf_checked\": %ld, \"logf_mismatches\": %ld}\n", n_tanh, bad_tanh, n_log, bad_log);
    return bad_tanh || bad_log || bad_atanh || bad_exp || bad_l1p;
}
