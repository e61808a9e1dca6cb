"""Region-level guard of the rational Genz families (integrands.h: box_safe): when it accepts a box, every point of
the box must give the in-range reciprocal an argument inside its exact range -- the scalar rule kernel then runs the
variant WITHOUT per-point checks (rule_kernels.cuh: CUB_GM_HAS_BOX_GUARD).  Host build of the same header; no GPU."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
DIM = 4
LAMBDAS = (0.0, 0.3585685828003181, 0.9486832980505138, 0.6882472016116853, 1.0)


@pytest.fixture(scope="module")
def bg():
    d = tempfile.mkdtemp(prefix="box_guard_")
    so = os.path.join(d, "libbg.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", so,
                           os.path.join(HERE, "box_guard_harness.cpp")])
    L = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    L.bg_check.restype = ctypes.c_int
    L.bg_check.argtypes = [ctypes.c_int, dp, dp, dp, dp, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    L.bg_exp_pair.restype = None
    L.bg_exp_pair.argtypes = [ctypes.c_double, dp]
    return L


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _points(rng, c, h, n):
    """Corners, rule-point offsets and random points of the box c +- h."""
    lam = np.array(LAMBDAS)
    x = [c + s * l * h for l in lam for s in (-1.0, 1.0)]
    signs = rng.choice([-1.0, 1.0], (n, DIM))
    x += list(c + signs * rng.choice(lam, (n, DIM)) * h)
    x += list(c + rng.uniform(-1, 1, (n, DIM)) * h)
    return np.ascontiguousarray(np.array(x))


def _check(bg, family, params, c, h, x):
    flag = ctypes.c_int(0)
    safe = bg.bg_check(family, _dp(params), _dp(c), _dp(h), _dp(x), len(x), ctypes.byref(flag))
    return bool(safe), bool(flag.value)


@pytest.mark.parametrize("family", [3, 2, 4, 5, 6, 10])
def test_box_safe_implies_every_argument_in_range(bg, family):
    rng = np.random.default_rng(5)
    accepted = rejected = 0
    for trial in range(4000):
        scale = 10.0 ** rng.uniform(-90, 90) if trial % 3 else 10.0 ** rng.uniform(-2, 2)
        if family == 3:
            a = rng.uniform(-1, 1, DIM) * scale
            if trial % 5 == 0:
                a = np.abs(a)
            params = np.concatenate([a, np.zeros(DIM)])
        elif family in (4, 5, 6):  # exp families: a (any sign, scales around the guard's threshold), u
            a = rng.uniform(-1, 1, DIM) * 10.0 ** rng.uniform(-2, 3.5)
            params = np.concatenate([a, rng.uniform(0, 1, DIM)])
        elif family == 10:
            params = np.concatenate([[rng.uniform(-1, 1) * 10.0 ** rng.uniform(-2, 4)], np.zeros(2 * DIM - 1)])
        else:
            params = np.concatenate([10.0 ** rng.uniform(-1, 1, DIM) * (scale if trial % 2 else 1.0), rng.uniform(0, 1, DIM)])
        c = rng.uniform(-1, 2, DIM)
        h = 10.0 ** rng.uniform(-12, 0.3, DIM)
        safe, in_range = _check(bg, family, params, c, h, _points(rng, c, h, 64))
        if safe:
            accepted += 1
            assert in_range, (family, params, c, h)
        else:
            rejected += 1
    assert accepted > 200 and rejected > 200  # both branches are exercised


@pytest.mark.parametrize("family", [3, 2, 4, 5, 6])
def test_box_safe_accepts_the_benchmark_boxes(bg, family):
    """Genz parameters as the benchmarks use them on sub-boxes of [0,1]^d: the guard must not send them to the slow path."""
    rng = np.random.default_rng(6)
    for _ in range(500):
        a = rng.uniform(0.1, 3.0, DIM)
        params = np.concatenate([a, rng.uniform(0, 1, DIM)])
        h = 0.5 * 2.0 ** -rng.integers(0, 30, DIM).astype(float)
        c = np.clip(rng.uniform(0, 1, DIM), h, 1 - h)
        safe, in_range = _check(bg, family, params, c, h, _points(rng, c, h, 8))
        assert safe and in_range


@pytest.mark.parametrize("family", [3, 2])
def test_box_safe_rejects_nan_zero_and_overflow(bg, family):
    rng = np.random.default_rng(7)
    c = np.full(DIM, 0.5)
    h = np.full(DIM, 0.25)
    x = _points(rng, c, h, 4)
    if family == 3:
        bad = [np.array([-0.5, -0.5, -0.5, -0.5, 0, 0, 0, 0.0]),  # s crosses zero inside the box
               np.array([1e70, 1e70, 1e70, 1e70, 0, 0, 0, 0.0]),   # s^5 overflows
               np.array([np.nan, 1, 1, 1, 0, 0, 0, 0.0]), np.array([np.inf, 1, 1, 1, 0, 0, 0, 0.0])]
    else:
        bad = [np.array([1e-80, 1e-80, 1e-80, 1e-80, 0.5, 0.5, 0.5, 0.5]),  # a^-2 huge: the product overflows
               np.array([1e200, 1.0, 1.0, 1.0, 0.5, 0.5, 0.5, 0.5]),         # a^-2 underflows to 0: a factor can vanish
               np.array([np.nan, 1, 1, 1, 0.5, 0.5, 0.5, 0.5]), np.array([1.0, 1, 1, 1, np.nan, 0.5, 0.5, 0.5])]
    for p in bad:
        safe, _ = _check(bg, family, p, c, h, x)
        assert not safe, p
    safe, _ = _check(bg, family, np.array([1.0, 1, 1, 1, 0.5, 0.5, 0.5, 0.5]), np.array([np.nan, 0.5, 0.5, 0.5]), h, x)
    assert not safe


def test_exp_inrange_has_the_bits_of_exp(bg):
    """dm_exp_inrange (no out-of-range selects, one scaling step) == dm_exp bit for bit on [-700, 700]."""
    rng = np.random.default_rng(8)
    xs = np.concatenate([rng.uniform(-700, 700, 200000), rng.uniform(-1, 1, 50000), 10.0 ** rng.uniform(-300, 2.8, 20000),
                         -(10.0 ** rng.uniform(-300, 2.8, 20000)), [0.0, -0.0, 700.0, -700.0, 690.0, -690.0, 1e-320, -1e-320]])
    out = np.zeros(2)
    for x in xs:
        bg.bg_exp_pair(float(x), _dp(out))
        assert out.view(np.uint64)[0] == out.view(np.uint64)[1], x
