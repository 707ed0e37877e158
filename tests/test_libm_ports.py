"""The f32 Phi and Tanh rules run bit-exact ports of glibc's tanhf / logf / atanhf (ldpc_toolbox_b200/csrc/libm_exact.h) so that
`Phif32` / `HLPhif32` reproduce a reference running on the platform libm.  This compiles the SAME header for the host
(plain operators, -ffp-contract=off) and compares it with the system libm on every float of the domain of use:
2 x 880 803 841 arguments of tanhf in +-[2^-100, 32], 1 115 684 864 of logf in (0, 64] and 2 x 1 065 353 217 of atanhf
in [-1, 1] (the f32 Tanh rule).  CPU only (~20 s on 8 cores)."""
import json
import os
import subprocess

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_libm_exact_header_matches_system_libm(tmp_path):
    exe = tmp_path / "libm_port_check"
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", os.path.join(ROOT, "tests", "libm_port_check.c"), "-o", str(exe), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=900)
    res = {}
    for line in out.stdout.strip().splitlines():
        res.update(json.loads(line))
    assert res["tanhf_checked"] == 880803841 and res["logf_checked"] == 1115684864 and res["atanhf_checked"] == 1065353217
    assert res["tanhf_mismatches"] == 0 and res["logf_mismatches"] == 0 and res["atanhf_mismatches"] == 0, res
