"""Seeded integration problems shared by the oracle tests, the GPU parity tests, smoke() and bench.py."""
import numpy as np

NAN = float("nan")

GENZ = {"oscillatory": 1, "product_peak": 2, "corner_peak": 3, "gaussian": 4, "c0": 5, "discontinuous": 6}
# Genz difficulty ||a||_1 at d = 5
GENZ_DIFFICULTY = {"oscillatory": 9.84, "product_peak": 24.0, "corner_peak": 24.0, "gaussian": 20.0, "c0": 6.0, "discontinuous": 4.0}


def genz_params(name, dim, seed=20260101, difficulty=None):
    """a ~ U(0,1) rescaled to the family's difficulty, u ~ U(0,1); params = a then u."""
    rng = np.random.default_rng(seed + 1000 * GENZ[name] + dim)
    a = rng.uniform(0.0, 1.0, dim)
    u = rng.uniform(0.0, 1.0, dim)
    diff = GENZ_DIFFICULTY[name] if difficulty is None else difficulty
    a = a * (diff / a.sum())
    return np.concatenate([a, u])


def genz_exact(name, dim, params):
    """Closed forms of SURVEY.md Appendix D, evaluated with mpmath."""
    import mpmath as mp
    mp.mp.dps = 40
    a = [mp.mpf(float(v)) for v in params[:dim]]
    u = [mp.mpf(float(v)) for v in params[dim:]]
    if name == "oscillatory":
        r = mp.cos(2 * mp.pi * u[0] + sum(a) / 2)
        for ai in a:
            r *= 2 * mp.sin(ai / 2) / ai
        return float(r)
    if name == "product_peak":
        r = mp.mpf(1)
        for ai, ui in zip(a, u):
            r *= ai * (mp.atan(ai * (1 - ui)) + mp.atan(ai * ui))
        return float(r)
    if name == "corner_peak":
        s = mp.mpf(0)
        for mask in range(1 << dim):
            t = mp.mpf(1)
            bits = 0
            for i in range(dim):
                if (mask >> i) & 1:
                    t += a[i]
                    bits += 1
            s += (-1) ** bits / t
        r = s / mp.factorial(dim)
        for ai in a:
            r /= ai
        return float(r)
    if name == "gaussian":
        r = mp.mpf(1)
        for ai, ui in zip(a, u):
            r *= mp.sqrt(mp.pi) / (2 * ai) * (mp.erf(ai * (1 - ui)) + mp.erf(ai * ui))
        return float(r)
    if name == "c0":
        r = mp.mpf(1)
        for ai, ui in zip(a, u):
            r *= (2 - mp.exp(-ai * ui) - mp.exp(-ai * (1 - ui))) / ai
        return float(r)
    if name == "discontinuous":
        r = mp.mpf(1)
        for i, (ai, ui) in enumerate(zip(a, u)):
            r *= (mp.exp(ai * ui) - 1) / ai if i < 2 else (mp.exp(ai) - 1) / ai
        return float(r)
    raise ValueError(name)


def random_boxes(dim, n, seed, lo=0.0, hi=1.0):
    rng = np.random.default_rng(seed)
    a = rng.uniform(lo, hi, (n, dim))
    b = rng.uniform(lo, hi, (n, dim))
    c = 0.5 * (a + b)
    h = 0.5 * np.abs(a - b) + 1e-3
    return c, h


def ulp_diff(a, b):
    """max distance in units in the last place between two float64 arrays (0 = bit-identical)."""
    a = np.ascontiguousarray(a, dtype=np.float64).ravel()
    b = np.ascontiguousarray(b, dtype=np.float64).ravel()
    ia = a.view(np.int64).copy()
    ib = b.view(np.int64).copy()
    ia[ia < 0] = np.int64(-2**63) - ia[ia < 0]
    ib[ib < 0] = np.int64(-2**63) - ib[ib < 0]
    return int(np.max(np.abs(ia - ib))) if a.size else 0


def bit_equal(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))
