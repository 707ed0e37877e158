#!/usr/bin/env python3
"""One-off extraction of the *standards tables* (numbers only) that define the
DVB-S2, 5G-NR and CCSDS AR4JA parity-check matrices.

The numbers are facts published in ETSI EN 302 307-1 Annex B/C, 3GPP TS 38.212
Tables 5.3.2-2/3 and CCSDS 131.0-B-5 Tables 7-3/7-4.  There is no network in
the build container, so they are read here from the literals in the reference
checkout (src/codes/dvbs2.rs:205-2170, src/codes/nr5g.rs:348-1134,
src/codes/ccsds.rs:230-420) and re-emitted in a compact whitespace format of
our own under ldpc_toolbox_b200/codes/data/.  Only this script ever touches
/root/reference; the generated data files are committed and are what
ldpc_toolbox_b200.codes reads at run time.

Usage:  python tools/extract_standard_tables.py [/root/reference]
"""
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(__file__), "..", "ldpc_toolbox_b200", "codes", "data")


def dvbs2():
    src = open(os.path.join(REF, "src/codes/dvbs2.rs")).read()
    # q values
    qsec = src[src.index("const fn q(self)"):src.index("const fn addresses(self)")]
    q = {m.group(1): int(m.group(2)) for m in re.finditer(r"Code::(\w+) => (\d+),", qsec)}
    asec = src[src.index("const fn addresses(self)"):]
    asec = asec[:asec.index("#[cfg(test)]")]
    parts = re.split(r"Code::(\w+) => &\[", asec)[1:]
    lines = []
    for name, body in zip(parts[0::2], parts[1::2]):
        rows = re.findall(r"&\[([^\]]*)\]", body)
        n = 16200 if name.endswith("short") else 64800
        lines.append(f"code {name} n {n} q {q[name]} rows {len(rows)}")
        for r in rows:
            nums = [int(x) for x in re.findall(r"\d+", r)]
            lines.append(" ".join(map(str, nums)))
    open(os.path.join(OUT, "dvbs2_addresses.txt"), "w").write("\n".join(lines) + "\n")


def nr5g():
    src = open(os.path.join(REF, "src/codes/nr5g.rs")).read()
    out = []
    for bg, fn in (("1", "fn base_graph_1()"), ("2", "fn base_graph_2()")):
        sec = src[src.index(fn):]
        nxt = sec.find("\nfn ", 10)
        sec = sec[:nxt] if nxt > 0 else sec
        vecs = re.split(r"vec!\[", sec)[1:]
        out.append(f"basegraph {bg} rows {len(vecs)}")
        for i, v in enumerate(vecs):
            rows = re.findall(r"row!\(([^)]*)\)", v)
            out.append(f"row {i} entries {len(rows)}")
            for r in rows:
                out.append(" ".join(r.split()))
    open(os.path.join(OUT, "nr5g_basegraphs.txt"), "w").write("\n".join(out) + "\n")


def ccsds():
    src = open(os.path.join(REF, "src/codes/ccsds.rs")).read()
    m = re.search(r"THETA_K: \[u8; 26\] = \[([^\]]*)\]", src)
    theta = [int(x) for x in re.findall(r"\d+", m.group(1))]
    psec = src[src.index("PHI_K:"):]
    psec = psec[psec.index("= [") :]
    psec = psec[: psec.index("];") + 2]
    # 4 blocks (j) x 26 rows (k) x 7 (M index)
    rows = re.findall(r"\[((?:\s*\d+\s*,?){7})\]", psec)
    assert len(rows) == 4 * 26, len(rows)
    out = ["theta " + " ".join(map(str, theta))]
    for j in range(4):
        for k in range(26):
            nums = [int(x) for x in re.findall(r"\d+", rows[j * 26 + k])]
            out.append(f"phi {j} {k + 1} " + " ".join(map(str, nums)))
    open(os.path.join(OUT, "ccsds_ar4ja.txt"), "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    dvbs2()
    nr5g()
    try:
        ccsds()
    except Exception as e:  # table layout differs: report, keep the others
        print("ccsds extraction failed:", e)
    print("tables written to", os.path.abspath(OUT))
