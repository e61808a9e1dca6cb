"""The big-integer arithmetic behind the exact running sums (cubature_b200/csrc/exact_acc.h), checked against Python
integers: a sum of doubles converted back truncated toward zero / rounded to nearest even, and the (H, L1, Ldbl) split the
select kernel uses for the residual inside a binade.  Host-only (g++); the same header is compiled for the device."""
import ctypes
import os
import struct
import subprocess
import tempfile
from fractions import Fraction

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def xa():
    d = tempfile.mkdtemp(prefix="xa_harness_")
    so = os.path.join(d, "libxa.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", so, os.path.join(HERE, "xa_harness.cpp")])
    L = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    L.xa_sum.restype = ctypes.c_double
    L.xa_sum.argtypes = [dp, ctypes.c_int, ctypes.c_int]
    L.xa_diff.restype = ctypes.c_double
    L.xa_diff.argtypes = [dp, ctypes.c_int, dp, ctypes.c_int, ctypes.c_int]
    L.xa_residual.restype = ctypes.c_double
    L.xa_residual.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
    L.xa_window_sum.restype = ctypes.c_double
    L.xa_window_sum.argtypes = [dp, ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    return L


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def to_double(fr, nearest):
    """Fraction >= 0 -> double, truncated toward zero or rounded to nearest even (reference implementation)."""
    if fr == 0:
        return 0.0
    n = fr * (1 << 1074)  # units of the smallest subnormal
    assert n.denominator == 1
    n = n.numerator
    t = n.bit_length() - 1
    if t <= 52:
        return struct.unpack("<d", struct.pack("<Q", n))[0]
    sh = t - 52
    m, rest = n >> sh, n & ((1 << sh) - 1)
    if nearest:
        half = 1 << (sh - 1)
        if rest > half or (rest == half and (m & 1)):
            m += 1
            if m >> 53:
                m >>= 1
                sh += 1
    e = sh + 1
    if e >= 2047:
        return float("inf")
    return struct.unpack("<d", struct.pack("<Q", (e << 52) | (m & ((1 << 52) - 1))))[0]


def samples(rng, n, kind):
    if kind == "wide":
        return np.ldexp(rng.uniform(0.5, 1.0, n), rng.integers(-1070, 900, n))
    if kind == "narrow":
        return rng.uniform(0.0, 1.0, n) * 1e-9
    if kind == "subnormal":
        return np.ldexp(rng.uniform(0.0, 1.0, n), rng.integers(-1074, -1000, n))
    if kind == "cancel":  # many equal values plus a tail far below: long carry chains
        a = np.full(n, 1.0 - 2.0 ** -53)
        a[::7] = 2.0 ** -200
        return a
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["wide", "narrow", "subnormal", "cancel"])
def test_sum_truncated_and_nearest(xa, kind):
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 17, 1000):
        for _ in range(20):
            x = samples(rng, n, kind)
            exact = sum(Fraction(float(v)) for v in x)
            for nearest in (0, 1):
                got = xa.xa_sum(_dp(x), n, nearest)
                assert got == to_double(exact, nearest), (kind, n, nearest, got.hex(), to_double(exact, nearest).hex())


def test_truncation_makes_less_than_exact(xa):
    """`double < tol` on the truncated sum must be the exact comparison `sum < tol`."""
    rng = np.random.default_rng(6)
    for _ in range(300):
        x = rng.uniform(0.0, 1.0, 50)
        exact = sum(Fraction(float(v)) for v in x)
        t = xa.xa_sum(_dp(x), 50, 0)
        for tol in (t, np.nextafter(t, np.inf), np.nextafter(t, 0.0)):
            assert (t < tol) == (exact < Fraction(float(tol)))


def test_difference(xa):
    rng = np.random.default_rng(7)
    for _ in range(200):
        a = samples(rng, 40, "wide")
        k = int(rng.integers(0, 40))
        b = a[rng.permutation(40)[:k]].copy()
        exact = sum(Fraction(float(v)) for v in a) - sum(Fraction(float(v)) for v in b)
        for nearest in (0, 1):
            assert xa.xa_diff(_dp(a), 40, _dp(b), k, nearest) == to_double(exact, nearest)


def test_residual_inside_a_binade(xa):
    """Residual after popping regions of one binade, through the (H, L1, Ldbl) split of k_exact_select."""
    rng = np.random.default_rng(8)
    for trial in range(400):
        binade = int(rng.integers(0, 2040)) if trial % 4 else int(rng.integers(0, 70))
        n_top = int(rng.integers(1, 30))
        if binade == 0:
            m = rng.integers(0, 1 << 52, n_top, dtype=np.uint64)
        else:
            m = rng.integers(1 << 52, 1 << 53, n_top, dtype=np.uint64)
        top = np.array([struct.unpack("<d", struct.pack("<Q", (binade << 52) | (int(v) & ((1 << 52) - 1))))[0] for v in m])
        # regions below the binade (possibly none, possibly far below)
        n_low = int(rng.integers(0, 20))
        lows = []
        for _ in range(n_low):
            e = int(rng.integers(0, binade + 1)) if binade > 0 else 0
            e = max(0, e - int(rng.integers(0, 3)))
            if e >= binade:
                e = max(0, binade - 1)
            fr = int(rng.integers(0, 1 << 52))
            lows.append(struct.unpack("<d", struct.pack("<Q", (e << 52) | fr))[0] if binade > 0 else 0.0)
        x = np.concatenate([top, np.array(lows, dtype=np.float64)]) if lows else top
        k = int(rng.integers(0, n_top + 1))
        popped = np.ascontiguousarray(m[:k])
        exact = sum(Fraction(float(v)) for v in x) - sum(Fraction(float(v)) for v in top[:k])
        got = xa.xa_residual(_dp(np.ascontiguousarray(x)), len(x), binade, popped.ctypes.data_as(ctypes.POINTER(ctypes.c_ulonglong)), k)
        assert got == to_double(exact, 0), (binade, k, n_top, n_low, got.hex(), to_double(exact, 0).hex())


def test_thread_windows_and_cells(xa):
    """The rule kernels' accumulation (acc_window.h): signed values spread over T per-thread 256-bit windows at some base,
    outliers through the far cells, windows summed as 32-bit cells and folded like the decide kernel does -- equal to the
    exact signed sum for every base, including windows that are wrong for the data."""
    rng = np.random.default_rng(9)
    for trial in range(300):
        n = int(rng.integers(1, 400))
        T = int(rng.choice([1, 32, 128]))
        e0 = int(rng.integers(-1000, 900))
        spread = int(rng.choice([4, 40, 150, 400]))
        x = np.ldexp(rng.uniform(0.5, 1.0, n), e0 + rng.integers(-spread, 1, n))
        x *= rng.choice([-1.0, 1.0], n)
        if trial % 5 == 0:
            x[rng.integers(0, n)] = 0.0
        if trial % 7 == 0:
            x[: n // 2] = np.ldexp(rng.uniform(0.0, 1.0, n // 2), -1074 + rng.integers(0, 60, n // 2))  # subnormals
        sgn = rng.choice([-1, 1], n).astype(np.int32)
        # window base: where the iteration kernel would put it (top of the data 8 bits below the window top), or off
        top = int(np.max(np.frexp(np.abs(x[x != 0]))[1])) + 1022 - 1 if np.any(x != 0) else 0  # biased exponent - 1 = lsb pos
        base = max(0, top + 8 - 159)
        if trial % 3 == 1:
            base = int(rng.integers(0, 1900))
        exact = sum(Fraction(float(v)) * int(s) for v, s in zip(x, sgn))
        n_far = ctypes.c_int(0)
        for nearest in (0, 1):
            got = xa.xa_window_sum(_dp(x), sgn.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), n, T, base, nearest, ctypes.byref(n_far))
            want = to_double(abs(exact), nearest) * (-1.0 if exact < 0 else 1.0)
            assert got == want, (trial, n, T, base, nearest, got.hex(), want.hex(), n_far.value)
