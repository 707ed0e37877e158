"""Exhaustive identity checks for the arithmetic shortcuts of the int8 flooding kernel
(ldpc_toolbox_b200/csrc/flood_i8.cu) against the direct form of the reference:

  g(a, acc) = max(0, min(a, acc) - T[|a - acc|])            reference src/decoder/arithmetic.rs:741
  h(a, acc) = max(0, min(a, acc) - T[|a - acc|] + T[sat_add(a, acc)])     arithmetic.rs:1155-1157
  T[t]      = round(8 ln(1 + e^{-t/8})) while positive       arithmetic.rs:589-598

The kernel never evaluates g directly.  It keeps a fold chain as the index of its next table read,
idx = a_next - acc, reads V[idx] = max(idx, 0) + T[|idx|] and forms the next index as
min(V[idx] + (a' - a), a'); a chain ends with acc = max(a - V[idx], 0).  The A-Min* path and the
generic-degree path use U[d] = min(d, 0) - T[|d|].  Every identity is checked on the whole domain
(magnitudes 0..127).  Also pins the prefix-sharing order of a whole check node (25 fold steps at
degree 7) and the packed variable node against scalar restatements.  CPU only.
"""
import math

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import pyref


def table_T():
    t = []
    for i in range(128):
        x = round(8.0 * math.log1p(math.exp(-(i / 8.0))))
        if x <= 0:
            break
        t.append(x)
    full = np.zeros(256, dtype=np.int64)
    full[:len(t)] = t
    return full


T = table_T()
A = np.arange(128, dtype=np.int64)


def g_direct(a, acc):
    return np.maximum(0, np.minimum(a, acc) - T[np.abs(a - acc)])


def h_direct(a, acc):
    return np.maximum(0, np.minimum(a, acc) - T[np.abs(a - acc)] + T[np.minimum(a + acc, 127)])


def U(d):
    return np.minimum(d, 0) - T[np.abs(d)]


def V(d):
    return np.maximum(d, 0) + T[np.abs(d)]


def test_table_is_the_threshold_count():
    # flood_i8.cu table_T(): T[t] = #{theta in {1,3,5,9,13,22} : t < theta}; matches pyref's table too
    thetas = (1, 3, 5, 9, 13, 22)
    for t in range(256):
        assert T[t] == sum(t < th for th in thetas)
    assert [pyref._lookup(t) for t in range(128)] == T[:128].tolist()


def test_u_form_exhaustive():
    a, acc = np.meshgrid(A, A, indexing="ij")
    assert (np.maximum(0, acc + U(a - acc)) == g_direct(a, acc)).all()
    assert (np.maximum(0, acc + U(a - acc) + T[np.minimum(a + acc, 127)]) == h_direct(a, acc)).all()


def test_v_tables_fit_their_storage():
    d = np.arange(-127, 128)
    assert U(d).min() >= -128 and U(d).max() <= 127          # int8_t U[256]
    assert V(d).min() >= 0 and V(d).max() <= 255             # uint8_t V[256]
    assert (V(d) == d - U(d)).all()


def test_difference_form_step_exhaustive():
    """One fold step in difference form, all (a, acc, a') in [0,127]^3."""
    a, acc, an = np.meshgrid(A, A, A, indexing="ij")
    idx = a - acc
    new_acc = g_direct(a, acc)
    nxt = np.minimum(V(idx) + (an - a), an)                  # kernel: __viaddmin_s32(v, d1, a')
    assert (nxt == an - new_acc).all()
    assert (nxt >= -127).all() and (nxt <= 127).all()        # stays inside the table
    a2, acc2 = np.meshgrid(A, A, indexing="ij")
    assert (np.maximum(a2 - V(a2 - acc2), 0) == g_direct(a2, acc2)).all()     # chain end: __viaddmax_s32_relu(a, -v, 0)


def check_node_direct(x):
    """arithmetic.rs:722-751 on magnitudes: for every excluded j, fold the others left to right."""
    out = []
    for j in range(len(x)):
        acc = None
        for i, v in enumerate(x):
            if i == j:
                continue
            acc = v if acc is None else int(g_direct(np.int64(v), np.int64(acc)))
        out.append(acc)
    return out


def check_node_kernel_order(x):
    """flood_i8.cu check_word (non-AMIN, D >= 3): shared prefix P_j, chains in difference form."""
    D = len(x)
    d1 = [x[i + 1] - x[i] for i in range(D - 1)]
    steps = 0

    def run(idx, i0):
        nonlocal steps
        out = 0
        for i in range(2, D):
            if i < i0:
                continue
            v = int(V(np.int64(idx)))
            steps += 1
            if i + 1 < D:
                idx = min(v + d1[i], x[i + 1])
            else:
                out = max(x[i] - v, 0)
        return out

    r = [0] * D
    r[0] = run(d1[1], 2)
    r[1] = run(x[2] - x[0], 2)
    dP = d1[0]
    for j in range(2, D):
        v = int(V(np.int64(dP)))
        steps += 1
        if j == D - 1:
            r[j] = max(x[j - 1] - v, 0)
        else:
            dP = min(v + d1[j - 1], x[j])
            r[j] = run(dP + d1[j], j + 1)
    return r, steps


@settings(max_examples=400, deadline=None)
@given(st.lists(st.integers(0, 127), min_size=3, max_size=10))
def test_check_node_prefix_sharing(x):
    r, steps = check_node_kernel_order(x)
    assert r == check_node_direct(x)
    D = len(x)
    assert steps == 2 * (D - 2) + (D - 2) + (D - 2) * (D - 3) // 2        # 25 at D = 7 (35 unshared)


def test_check_node_degree7_step_count():
    _, steps = check_node_kernel_order([5, 9, 1, 127, 0, 33, 64])
    assert steps == 25


@settings(max_examples=300, deadline=None)
@given(st.lists(st.integers(-127, 127), min_size=2, max_size=30), st.booleans())
def test_pyref_vs_cpp_random_rows(oracle, x, amin):
    """One check row of degree 2..30 over degree-1 variables: after one flooding iteration the posterior
    of variable j is x_j + c2v_j, so the two restatements' check-node rules meet on arbitrary inputs."""
    d = len(x)
    alist = f"{d} 1\n1 {d}\n" + " ".join(["1"] * d) + f"\n{d}\n" + "1\n" * d + " ".join(str(i + 1) for i in range(d)) + "\n"
    impl = "Aminstari8" if amin else "Minstarapproxi8"
    llrs = np.array(x, dtype=np.float64) / 8.0
    dec = oracle.decoder(alist, impl)
    out, it = dec.decode(llrs, 1)
    h = pyref.SparseMatrix.from_alist(alist)
    cw, pit, ok, ppost = pyref.decode(h, impl, llrs.tolist(), 1)
    assert out.tolist() == cw and it == (pit if ok else -1)
    if it != 0:
        assert dec.posteriors().tolist() == [float(v) for v in ppost]


def var_node_direct(inp, c, jones, deg1clip):
    """arithmetic.rs:622-654 (+ :806-810 Jones, :826-842 degree-one clip)."""
    clip = lambda v: max(-127, min(127, v))
    i0 = inp
    if deg1clip and len(c) == 1:
        i0 = max(-116, min(116, inp))
    llr = i0 + sum(c)
    if jones:
        llr = clip(llr)
    return [clip(llr - v) for v in c], clip(llr)


def var_node_kernel(inp, c, jones, deg1clip):
    """flood_i8.cu var_class on one 16-bit half: offset-binary bytes, biased sum, VIADDMNMX.S16x2 clamp."""
    d = len(c)
    s = inp + 128
    if deg1clip and d == 1:
        s = max(12, min(244, s))
    for v in c:
        s += v + 128
    L = s - 128 * (d + 1)
    if jones:
        L = max(-127, min(127, L))
        base, negK = L + 384, -256 + 128
    else:
        base, negK = s, -128 * d + 128
    hard = 1 if (L - 1) < 0 else 0
    out = [max(min(base - (v + 128) + negK, 255), 1) - 128 for v in c]
    return out, hard


@settings(max_examples=500, deadline=None)
@given(st.integers(-127, 127), st.lists(st.integers(-127, 127), min_size=1, max_size=13), st.booleans(), st.booleans())
def test_packed_variable_node(inp, c, jones, deg1clip):
    ref_out, ref_llr = var_node_direct(inp, c, jones, deg1clip)
    out, hard = var_node_kernel(inp, c, jones, deg1clip)
    assert out == ref_out
    assert hard == (1 if ref_llr <= 0 else 0)
