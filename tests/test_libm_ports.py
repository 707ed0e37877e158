"""The f32 Phi and Tanh rules run bit-exact ports of glibc's tanhf / logf / atanhf (ldpc_toolbox_b200/csrc/libm_exact.h) so that
`Phif32` / `HLPhif32` reproduce a reference running on the platform libm.  This compiles the SAME header for the host
(plain operators, -ffp-contract=off) and compares it with the system libm on every float of the domain of use:
2 x 880 803 841 arguments of tanhf in +-[2^-100, 32], 1 115 684 864 of logf in (0, 64] and 2 x 1 065 353 217 of atanhf
in [-1, 1] (the f32 Tanh rule), 2 x 1 120 927 745 of expf in [-104, 104] and 1 073 741 825 of log1pf in [0, 2] (the opt-in
exact mode of the f32 min* rules).  CPU only (~25 s on 8 cores; needs an FMA-capable x86-64 CPU, as glibc's dispatch does)."""
import json
import os
import subprocess

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_libm_exact_header_matches_system_libm(tmp_path):
    exe = tmp_path / "libm_port_check"
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-fopenmp", os.path.join(ROOT, "tests", "libm_port_check.c"), "-o", str(exe), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=900)
    res = {}
    for line in out.stdout.strip().splitlines():
        res.update(json.loads(line))
    assert res["tanhf_checked"] == 880803841 and res["logf_checked"] == 1115684864 and res["atanhf_checked"] == 1065353217
    assert res["tanhf_mismatches"] == 0 and res["logf_mismatches"] == 0 and res["atanhf_mismatches"] == 0, res
    # expf (x86-64 glibc's FMA build, which every FMA-capable CPU dispatches to) and log1pf: the exact ln(1 + e^-t) of
    # the opt-in bit-exact mode of the f32 Min*-approx / A-Min* rules (LDPC_B200_EXACT_LIBM=1)
    assert res["expf_checked"] == 1120927745 and res["log1pf_checked"] == 1073741825
    assert res["expf_mismatches"] == 0 and res["log1pf_mismatches"] == 0, res
