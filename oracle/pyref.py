"""oracle/pyref.py — TEST INFRASTRUCTURE ONLY.

A second, deliberately naive restatement of the reference decoders in pure
Python, written independently of oracle/ldpc_oracle.cpp and shaped like the Rust
(per-destination message lists located by linear search, decoder.rs:85-155), so
the two restatements can be played against each other on small codes
(tests/test_oracle_crosscheck.py).  f64 rules use Python floats (glibc libm, the
same libm the Rust `std` float methods call on Linux); f32 rules round every
intermediate through numpy.float32.

"parity unpinned": like the C++ oracle, only the Phif64 flooding path is pinned
by reference tests (src/decoder/flooding.rs:161-189).

Each function cites the reference file:line it restates.
"""
from __future__ import annotations

import math

import numpy as np


# ------------------------------------------------------------------ src/sparse.rs
class SparseMatrix:
    def __init__(self, nrows, ncols):
        self.rows = [[] for _ in range(nrows)]
        self.cols = [[] for _ in range(ncols)]

    def insert(self, r, c):                       # sparse.rs:114-119
        if r not in self.cols[c]:
            self.rows[r].append(c)
            self.cols[c].append(r)

    @staticmethod
    def from_alist(text):                         # sparse.rs:352-389
        lines = text.split("\n")
        ncols, nrows = (int(x) for x in lines[0].split()[:2])
        h = SparseMatrix(nrows, ncols)
        for col in range(ncols):
            for tok in lines[4 + col].split():
                row = int(tok)
                if row != 0:
                    h.insert(row - 1, col)
        return h


# ------------------------------------------------------------------ float helpers
class _F64:
    name = "f64"

    @staticmethod
    def c(x):
        return float(x)

    tanh = staticmethod(math.tanh)
    log = staticmethod(math.log)
    exp = staticmethod(math.exp)
    log1p = staticmethod(math.log1p)
    atanh = staticmethod(math.atanh)
    tanh_clamp = 18.0


class _F32:
    name = "f32"

    @staticmethod
    def c(x):
        return np.float32(x)

    tanh = staticmethod(lambda x: np.float32(np.tanh(np.float32(x))))
    log = staticmethod(lambda x: np.float32(np.log(np.float32(x))))
    exp = staticmethod(lambda x: np.float32(np.exp(np.float32(x))))
    log1p = staticmethod(lambda x: np.float32(np.log1p(np.float32(x))))
    atanh = staticmethod(lambda x: np.float32(np.arctanh(np.float32(x))))
    tanh_clamp = np.float32(9.0)


def _fmax(a, b):   # Rust f64::max ignores a NaN operand
    if a != a:
        return b
    if b != b:
        return a
    return a if a > b else b


def _fmin(a, b):
    if a != a:
        return b
    if b != b:
        return a
    return a if a < b else b


# ------------------------------------------------------------------ arithmetics
class FloatArith:
    """Common float behaviour, arithmetic.rs:140-156 and the trivial conversions."""

    def __init__(self, F):
        self.F = F

    def quantize(self, llr):
        return self.F.c(llr)

    def hard(self, llr):
        return llr <= 0

    def llr_to_var_llr(self, l):
        return l

    def var_llr_to_llr(self, v):
        return v

    def send_var(self, input_llr, msgs):          # :140-156
        s = self.F.c(0.0)
        for m in msgs:
            s = self.F.c(s + m)
        llr = self.F.c(input_llr + s)
        return llr, [self.F.c(llr - m) for m in msgs]


class Phi(FloatArith):                            # arithmetic.rs:158-298
    def phi(self, x):
        F = self.F
        x = _fmax(x, F.c(1e-30))
        return F.c(-F.log(F.tanh(F.c(F.c(0.5) * x))))

    def send_check(self, xs):
        F = self.F
        sign, s, phis = 0, F.c(0.0), []
        for x in xs:
            p = self.phi(abs(x))
            phis.append(p)
            s = F.c(s + p)
            if x < 0:
                sign ^= 1
        out = []
        for x, p in zip(xs, phis):
            y = self.phi(F.c(s - p))
            sg = sign ^ 1 if x < 0 else sign
            out.append(y if sg == 0 else F.c(-y))
        return out

    def update_row(self, rcv, dest, vars_):
        F = self.F
        xs = [F.c(vars_[d] - r) for d, r in zip(dest, rcv)]
        new = self.send_check(xs)
        for j, d in enumerate(dest):
            rcv[j] = new[j]
            vars_[d] = F.c(xs[j] + new[j])


class Tanh(FloatArith):                           # arithmetic.rs:300-435
    def _t(self, x):
        F = self.F
        h = F.c(F.c(0.5) * x)
        c = F.tanh_clamp
        h = -c if h < -c else (c if h > c else h)
        return F.tanh(h)

    def send_check(self, xs):
        F = self.F
        ts = [self._t(x) for x in xs]
        out = []
        for j in range(len(xs)):
            p = F.c(1.0)
            for i, t in enumerate(ts):
                if i != j:
                    p = F.c(p * t)
            out.append(F.c(F.c(2.0) * F.atanh(p)))
        return out

    def update_row(self, rcv, dest, vars_):
        F = self.F
        xs = [F.c(vars_[d] - r) for d, r in zip(dest, rcv)]
        new = self.send_check(xs)
        for j, d in enumerate(dest):
            vars_[d] = F.c(vars_[d] + F.c(new[j] - rcv[j]))
            rcv[j] = new[j]


class MinstarApproxF(FloatArith):                 # arithmetic.rs:437-580
    def g(self, x, y):
        F = self.F
        return _fmax(F.c(_fmin(x, y) - F.log1p(F.exp(F.c(-abs(F.c(x - y)))))), F.c(0.0))

    def send_check(self, xs):
        F = self.F
        out = []
        for j in range(len(xs)):
            sign, acc = 0, None
            for i, x in enumerate(xs):
                if i == j:
                    continue
                if x < 0:
                    sign ^= 1
                a = abs(x)
                acc = a if acc is None else self.g(a, acc)
            if acc is None:
                raise RuntimeError("only one variable message connected to check node")
            out.append(acc if sign == 0 else F.c(-acc))
        return out

    def update_row(self, rcv, dest, vars_):
        F = self.F
        xs = [F.c(vars_[d] - r) for d, r in zip(dest, rcv)]
        new = self.send_check(xs)
        for j, d in enumerate(dest):
            vars_[d] = F.c(vars_[d] + F.c(new[j] - rcv[j]))
            rcv[j] = new[j]


class AminstarF(FloatArith):                      # arithmetic.rs:899-1072
    def h(self, x, y):
        F = self.F
        a = F.c(_fmin(x, y) - F.log1p(F.exp(F.c(-abs(F.c(x - y))))))
        return F.c(a + F.log1p(F.exp(F.c(-F.c(x + y)))))

    def _core(self, xs):
        argmin = min(range(len(xs)), key=lambda i: abs(xs[i]))   # first minimum
        sign, delta = 0, None
        for j, x in enumerate(xs):
            if x < 0:
                sign ^= 1
            if j != argmin:
                a = abs(x)
                delta = a if delta is None else self.h(a, delta)
        if delta is None:
            raise RuntimeError("var_messages_empty")
        d2 = self.h(delta, abs(xs[argmin]))
        return argmin, sign, delta, d2

    def send_check(self, xs):
        F = self.F
        argmin, sign, d1, d2 = self._core(xs)
        out = []
        for j, x in enumerate(xs):
            mag = d1 if j == argmin else d2
            out.append(F.c(-mag) if (sign != 0) ^ (x < 0) else mag)
        return out

    def update_row(self, rcv, dest, vars_):
        F = self.F
        xs = [F.c(vars_[d] - r) for d, r in zip(dest, rcv)]
        new = self.send_check(xs)
        for j, d in enumerate(dest):
            rcv[j] = new[j]
            vars_[d] = F.c(xs[j] + new[j])


_TABLE = []
for _t in range(128):                             # arithmetic.rs:588-597
    _x = int(math.floor(8.0 * math.log1p(math.exp(-_t / 8.0)) + 0.5))
    if _x > 0:
        _TABLE.append(_x)
    else:
        break


def _lookup(x):
    assert x >= 0
    return _TABLE[x] if x < len(_TABLE) else 0


def _clip(x):
    return 127 if x >= 127 else (-127 if x <= -127 else x)


class I8Arith:                                    # arithmetic.rs:582-654 (+ variants :806-848)
    def __init__(self, jones=False, hardlimit=False, deg1clip=False):
        self.jones, self.hl, self.deg1 = jones, hardlimit, deg1clip

    def quantize(self, llr):                      # :690-699
        x = 8.0 * llr
        if x >= 127.0:
            return 127
        if x <= -127.0:
            return -127
        if x != x:
            return 0
        return int(math.copysign(math.floor(abs(x) + 0.5), x))   # round half away from zero

    def hard(self, llr):
        return llr <= 0

    def llr_to_var_llr(self, l):
        return l

    def var_llr_to_llr(self, v):
        return _clip(v)

    def hardlimit(self, x):                       # :812-824
        if not self.hl:
            return x
        return -127 if x <= -100 else (127 if x >= 100 else x)

    def send_var(self, input_llr, msgs):          # :622-654
        inp = input_llr
        if self.deg1 and len(msgs) == 1:
            inp = -116 if inp <= -116 else (116 if inp >= 116 else inp)
        llr = inp + sum(msgs)
        if self.jones:
            llr = _clip(llr)
        return _clip(llr), [_clip(llr - m) for m in msgs]


class MinstarApproxI8(I8Arith):                   # arithmetic.rs:656-804
    def send_check(self, xs):
        out = []
        for j in range(len(xs)):
            sign, acc = 0, None
            for i, x in enumerate(xs):
                if i == j:
                    continue
                if x < 0:
                    sign ^= 1
                a = abs(x)
                acc = a if acc is None else max(min(a, acc) - _lookup(abs(a - acc)), 0)
            if acc is None:
                raise RuntimeError("only one variable message connected to check node")
            out.append(self.hardlimit(acc if sign == 0 else -acc))
        return out

    def update_row(self, rcv, dest, vars_):       # :759-801
        xs = [_clip(vars_[d] - r) for d, r in zip(dest, rcv)]
        new = self.send_check(xs)
        for j, d in enumerate(dest):
            vars_[d] += new[j] - rcv[j]
            rcv[j] = new[j]


class AminstarI8(I8Arith):                        # arithmetic.rs:1074-1260
    @staticmethod
    def h(x, y):
        return max(min(x, y) - _lookup(abs(x - y)) + _lookup(min(x + y, 127)), 0)

    def _core(self, xs):
        argmin = min(range(len(xs)), key=lambda i: abs(xs[i]))
        sign, delta = 0, None
        for j, x in enumerate(xs):
            if x < 0:
                sign ^= 1
            if j != argmin:
                a = abs(x)
                delta = a if delta is None else self.h(a, delta)
        if delta is None:
            raise RuntimeError("var_messages_empty")
        d2 = self.h(delta, abs(xs[argmin]))
        return argmin, sign, self.hardlimit(delta), self.hardlimit(d2)

    def send_check(self, xs):
        argmin, sign, d1, d2 = self._core(xs)
        return [(-(d1 if j == argmin else d2)) if (sign != 0) ^ (x < 0) else (d1 if j == argmin else d2)
                for j, x in enumerate(xs)]

    def update_row(self, rcv, dest, vars_):       # :1197-1257
        xs = [_clip(vars_[d] - r) for d, r in zip(dest, rcv)]
        argmin, sign, d1, d2 = self._core(xs)
        msgmin_rcv = -d1 if (sign != 0) ^ (xs[argmin] < 0) else d1
        for j, d in enumerate(dest):
            x = vars_[d] - rcv[j]                 # unclipped
            if j == argmin:
                r = msgmin_rcv
            else:
                r = -d2 if (sign != 0) ^ (x < 0) else d2
            vars_[d] = x + r
            rcv[j] = r


# ------------------------------------------------------------------ factory.rs:240-277
def _arith_for(name):
    hl = name.startswith("HL")
    base = name[2:] if hl else name
    for prefix, cls in (("Minstarapproxi8", MinstarApproxI8), ("Aminstari8", AminstarI8)):
        if base.startswith(prefix):
            flags = base[len(prefix):]
            return hl, cls(jones="Jones" in flags, hardlimit="PartialHardLimit" in flags, deg1clip="Deg1Clip" in flags)
    F = _F64 if base.endswith("f64") else _F32
    rule = base[:-3]
    cls = {"Phi": Phi, "Tanh": Tanh, "Minstarapprox": MinstarApproxF, "Aminstar": AminstarF}[rule]
    return hl, cls(F)


# ------------------------------------------------------------------ decoders
def _check_llrs(h, llrs, hd):                     # decoder.rs:157-164
    return not any(sum(1 for c in row if hd(llrs[c])) % 2 == 1 for row in h.rows)


def decode(h: SparseMatrix, name: str, llrs, max_iterations: int):
    """Returns (codeword list, iterations, success, posterior list)."""
    layered, A = _arith_for(name)
    n = len(h.cols)
    assert len(llrs) == n
    raw = lambda x: x <= 0.0
    if _check_llrs(h, llrs, raw):
        return [int(raw(x)) for x in llrs], 0, True, list(llrs)
    if layered:                                   # horizontal_layered.rs:49-110
        q = [A.llr_to_var_llr(A.quantize(x)) for x in llrs]
        rcv = [[0 if isinstance(q[0], int) else type(q[0])(0.0)] * len(r) for r in h.rows]
        hd = lambda x: A.hard(A.var_llr_to_llr(x))
        for it in range(1, max_iterations + 1):
            for r, row in enumerate(h.rows):
                A.update_row(rcv[r], row, q)
            if _check_llrs(h, q, hd):
                return [int(hd(x)) for x in q], it, True, q
        return [int(hd(x)) for x in q], max_iterations, False, q
    # flooding.rs:51-125, messages kept per destination and located by search
    inp = [A.quantize(x) for x in llrs]
    out = list(inp)
    var_msgs = [[[v, None] for v in row] for row in h.rows]      # per check: [source var, value]
    chk_msgs = [[[c, None] for c in col] for col in h.cols]      # per var:   [source check, value]

    def send(store, source, dest, value):
        for m in store[dest]:
            if m[0] == source:
                m[1] = value
                return
        raise RuntimeError("message for source not found")

    for v in range(n):
        for c in h.cols[v]:
            send(var_msgs, v, c, inp[v])
    for it in range(1, max_iterations + 1):
        for c, msgs in enumerate(var_msgs):
            if not msgs and not isinstance(A, (AminstarF, AminstarI8)):
                continue
            vals = A.send_check([m[1] for m in msgs])
            for m, val in zip(msgs, vals):
                send(chk_msgs, c, m[0], val)
        for v, msgs in enumerate(chk_msgs):
            out[v], vals = A.send_var(inp[v], [m[1] for m in msgs])
            for m, val in zip(msgs, vals):
                send(var_msgs, v, m[0], val)
        if _check_llrs(h, out, A.hard):
            return [int(A.hard(x)) for x in out], it, True, out
    return [int(A.hard(x)) for x in out], max_iterations, False, out


# ---------------------------------------------------------------------------------------------
# 8PSK + bit interleaver (test infrastructure, like everything in oracle/)
# ---------------------------------------------------------------------------------------------
def interleave(x, columns: int, backwards: bool = False):
    """reference src/simulation/interleaving.rs:40-58: write the codeword by rows into a
    (columns x len/columns) matrix, read it by columns (rows of the transpose optionally reversed)."""
    import numpy as np
    x = np.asarray(x)
    assert x.size % columns == 0
    t = x.reshape(columns, x.size // columns).T
    if backwards:
        t = t[:, ::-1]
    return np.ascontiguousarray(t).reshape(-1)


def deinterleave(x, columns: int, backwards: bool = False):
    """reference src/simulation/interleaving.rs:64-85"""
    import numpy as np
    x = np.asarray(x)
    assert x.size % columns == 0
    t = x.reshape(x.size // columns, columns).T
    if backwards:
        t = t[::-1, :]
    return np.ascontiguousarray(t).reshape(-1)


_PSK8 = None


def _psk8_points():
    global _PSK8
    if _PSK8 is None:
        a = math.sqrt(0.5)
        # reference src/simulation/modulation.rs:166-178, keyed by (b0, b1, b2)
        _PSK8 = {(0, 0, 0): complex(a, a), (1, 0, 0): complex(0, 1), (1, 1, 0): complex(-a, a), (0, 1, 0): complex(-1, 0),
                 (0, 1, 1): complex(-a, -a), (1, 1, 1): complex(0, -1), (1, 0, 1): complex(a, -a), (0, 0, 1): complex(1, 0)}
    return _PSK8


def psk8_modulate(bits):
    """reference src/simulation/modulation.rs:189-203"""
    bits = [int(b) for b in bits]
    assert len(bits) % 3 == 0
    pts = _psk8_points()
    return [pts[(bits[i], bits[i + 1], bits[i + 2])] for i in range(0, len(bits), 3)]


def _maxstar(a, b):
    return max(a, b) + math.log1p(math.exp(-abs(a - b)))


def psk8_demodulate(symbols, noise_sigma: float):
    """reference src/simulation/modulation.rs:222-262: exact LLRs with max*, reduce order as written there"""
    pts = _psk8_points()
    scale = 1.0 / (noise_sigma * noise_sigma)
    out = []
    for s in symbols:
        s = s * scale
        d = {k: s.real * p.real + s.imag * p.imag for k, p in pts.items()}

        def red(keys):
            acc = d[keys[0]]
            for k in keys[1:]:
                acc = _maxstar(acc, d[k])
            return acc
        out.append(red([(0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1)]) - red([(1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1)]))
        out.append(red([(0, 0, 0), (0, 0, 1), (1, 0, 0), (1, 0, 1)]) - red([(0, 1, 0), (0, 1, 1), (1, 1, 0), (1, 1, 1)]))
        out.append(red([(0, 0, 0), (0, 1, 0), (1, 0, 0), (1, 1, 0)]) - red([(0, 0, 1), (0, 1, 1), (1, 0, 1), (1, 1, 1)]))
    return out
