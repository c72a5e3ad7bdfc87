"""The E-step's 2^x (gingr_b200/csrc/exp2_tab.cuh + gauss_exp2_tab_u of exp2_poly.cuh), checked on the CPU: the generated
table against an independent high-precision evaluation, and a BIT-EXACT emulation of the device routine (IEEE adds, FMAs
through exact rationals, the integer exponent insertion) against the exact value over the whole argument range -- the
accuracy claim of DESIGN.md for K1 without needing the device.  No GPU."""
import pathlib
import os
import re
import struct
from fractions import Fraction

import numpy as np
import pytest

mp = pytest.importorskip("mpmath")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NBIAS, KBIAS = 2048, 64


def _parse():
    src = pathlib.Path(os.path.join(ROOT, "gingr_b200", "csrc", "exp2_tab.cuh")).read_text()
    n = int(re.search(r"GAUSS_TAB_N = (\d+)", src).group(1))
    bits = int(re.search(r"GAUSS_TAB_BITS = (\d+)", src).group(1))
    rows = [(int(a, 16), int(b, 16)) for a, b in re.findall(r"\{0x([0-9a-f]{8})u, 0x([0-9a-f]{8})u\}", src)]
    coef = [float.fromhex(h) for h in re.findall(r"GAUSS_EXP2_A\d = (\S+);", src)]
    return n, bits, rows, coef


def _fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))           # one rounding, as the hardware FMA


def _dbl(hi, lo):
    return struct.unpack("<d", struct.pack("<II", lo & 0xFFFFFFFF, hi & 0xFFFFFFFF))[0]


def _device_exp2(u, n, bits, rows, coef):
    """gauss_exp2_tab_u(u): returns 2^64 * 2^(u / n) as the kernel computes it (0.0 below the underflow cut)."""
    shift = 6755399441055744.0 + float(NBIAS * n)
    tmp = u + shift
    lo = struct.unpack("<q", struct.pack("<d", tmp))[0] & 0xFFFFFFFF
    lo_signed = lo - (1 << 32) if lo & 0x80000000 else lo
    nf = tmp - shift
    r = u - nf
    k = lo & (n - 1)
    tlo, thi = rows[k]
    T = _dbl(((lo << (20 - bits)) + thi) & 0xFFFFFFFF, tlo)
    p = coef[3]
    p = _fma(p, r, coef[2])
    p = _fma(p, r, coef[1])
    p = _fma(p, r, coef[0])
    s = T * r
    v = _fma(s, p, T)
    return v if lo_signed >= 963 * n else 0.0


def test_table_entries_and_coefficients():
    n, bits, rows, coef = _parse()
    assert n == 256 and bits == 8 and len(rows) == n and len(coef) == 4
    mp.mp.dps = 50
    for k, (lo, hi_adj) in enumerate(rows):
        want = float(mp.mpf(2) ** (mp.mpf(k) / n))                   # correctly rounded 2^(k/n)
        hi = (hi_adj + (k << (20 - bits)) + ((NBIAS - KBIAS) << 20)) & 0xFFFFFFFF
        assert _dbl(hi, lo) == want, k
    fact = 1
    for m in range(1, 5):
        fact *= m
        assert coef[m - 1] == float((mp.log(2) / n) ** m / fact)


def test_bit_exact_emulation_meets_the_accuracy_claim():
    n, bits, rows, coef = _parse()
    mp.mp.dps = 50
    rng = np.random.default_rng(0)
    us = list(-rng.uniform(0.0, 1085.0 * n, 1500)) + list(-rng.uniform(0.0, 40.0 * n, 500))
    us += [0.0, -0.5, -1.0, -127.5, -128.0, -128.5, -255.999, -256.0, -1022.0 * n, -1074.0 * n, -1084.99 * n]
    worst = 0.0
    for u in us:
        got = _device_exp2(float(u), n, bits, rows, coef)
        want = mp.mpf(2) ** (mp.mpf(float(u)) / n + KBIAS)
        assert got > 0.0
        worst = max(worst, float(abs(mp.mpf(got) - want) / want))
    # table entry (0.5 ulp) + product / final FMA roundings + truncation 3.8e-17: below one ulp of the result overall
    assert worst < 2.3e-16, worst
    # the 2^64 bias keeps every representable kernel value NORMAL: K = 2^-1074 (the smallest subnormal) maps to 2^-1010
    assert _device_exp2(-1074.0 * n, n, bits, rows, coef) == 2.0 ** (-1074 + 64)
    # exact zero below the cut (n < -1085), as the reference's exp underflows to 0 there
    for u in (-1085.51 * n, -1086.0 * n, -5000.0 * n, -2.0 ** 30):
        assert _device_exp2(u, n, bits, rows, coef) == 0.0
    # monotone across table and exponent boundaries
    grid = [-(127.0 + j / 16.0) for j in range(0, 64)] + [-(255.0 + j / 16.0) for j in range(0, 48)]
    vals = [_device_exp2(u, n, bits, rows, coef) for u in grid]
    assert all(a >= b for a, b in zip(vals[:64], vals[1:64])) and all(a >= b for a, b in zip(vals[64:], vals[65:]))
