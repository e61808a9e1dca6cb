"""Accuracy of the deterministic math (cubature_b200/csrc/detmath.h, host build) against mpmath."""
import math

import mpmath as mp
import numpy as np
import pytest

mp.mp.prec = 200


def ulp_err(got, exact):
    exact_f = float(exact)
    if exact_f == 0.0:
        return abs(got)
    u = math.ulp(exact_f)
    return abs(mp.mpf(got) - exact) / u


@pytest.mark.parametrize("fn,ref,lo,hi", [
    ("orc_dm_exp", mp.exp, -700.0, 700.0),
    ("orc_dm_exp", mp.exp, -1.0, 1.0),
    ("orc_dm_sin", mp.sin, -50.0, 50.0),
    ("orc_dm_cos", mp.cos, -50.0, 50.0),
    ("orc_dm_sin", mp.sin, -1e4, 1e4),
    ("orc_dm_cos", mp.cos, -1e4, 1e4),
    ("orc_dm_log", mp.log, 1e-300, 1e300),
    ("orc_dm_log", mp.log, 0.5, 2.0),
])
def test_accuracy(oracle, fn, ref, lo, hi):
    L = oracle.lib()
    rng = np.random.default_rng(3)
    if fn == "orc_dm_log" and hi > 1e10:
        xs = np.exp(rng.uniform(math.log(lo), math.log(hi), 1500))
    else:
        xs = rng.uniform(lo, hi, 1500)
    worst = 0.0
    for x in xs:
        x = float(x)
        exact = ref(mp.mpf(x))
        if fn in ("orc_dm_sin", "orc_dm_cos") and abs(exact) < 1e-3:
            continue  # relative accuracy near zeros is limited by the 3-term reduction; absolute is fine
        worst = max(worst, float(ulp_err(getattr(L, fn)(x), exact)))
    assert worst <= 2.0, worst


def test_special_values(oracle):
    L = oracle.lib()
    assert L.orc_dm_exp(0.0) == 1.0
    assert L.orc_dm_exp(-1000.0) == 0.0
    assert L.orc_dm_exp(1000.0) == float("inf")
    assert math.isnan(L.orc_dm_exp(float("nan")))
    assert L.orc_dm_exp(-740.0) > 0.0  # gradual underflow
    assert L.orc_dm_cos(0.0) == 1.0 and L.orc_dm_sin(0.0) == 0.0
    assert math.isnan(L.orc_dm_sin(float("inf")))
    assert L.orc_dm_log(1.0) == 0.0
    assert L.orc_dm_log(0.0) == float("-inf")
    assert math.isnan(L.orc_dm_log(-1.0))
    assert L.orc_dm_pow(2.0, 0.5) == pytest.approx(math.sqrt(2.0), rel=4e-16)
