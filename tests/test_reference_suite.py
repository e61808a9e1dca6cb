"""Port of the reference's own test-suite (src/test/java/de/labathome/cubature/TestCubature.java) to the GPU
path, including the two tests that are commented out there.  Integrands 0..8 are the
device family REFTEST (cubature_b200/csrc/integrands.h:RefTest, same operation order as TestCubature.java:109-177).
"""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NAN = float("nan")
RADIUS = 0.50124145262344534123412  # TestCubature.java:39


def S(n):
    """Surface area of the n-dimensional unit hypersphere, TestCubature.java:182-202."""
    if n % 2 == 0:
        val = 2 * math.pi ** (n * 0.5)
        n //= 2
        fact = 1
        while n > 1:
            n -= 1
            fact *= n
        return val / fact
    val = (1 << (n // 2 + 1)) * math.pi ** (n // 2)
    fact = 1
    while n > 2:
        n -= 2
        fact *= n
    return val / fact


def exact_integral(which, xmax):
    """TestCubature.java:204-221."""
    dim = len(xmax)
    if which == 0:
        return float(np.prod(np.sin(xmax)))
    if which == 2:
        return S(dim) * (RADIUS * 0.5) ** dim / dim
    return 1.0


@pytest.mark.parametrize("dim", [1, 2, 3, 4])
def test_individual_integrands(cb, oracle, dim):
    """testIndividualIntegrands (TestCubature.java:282-351) with the dimension loop widened back to 1..4."""
    max_integrand = 8 if dim == 3 else 7
    for which in range(max_integrand + 1):
        ig = cb.Integrand.builtin(cb.Family.REFTEST, dim, 1, [which])
        cap = 3_000_000  # the 3-D / 4-D indicator (which == 2) does not converge in any sane budget (the reference only runs dim <= 2)
        r, st, _ = cb.Cubature.integrate_ex(ig, [0.0] * dim, [1.0] * dim, 1e-2, NAN, cb.CubatureError.INDIVIDUAL, cap)
        o = oracle.integrate(20, dim, 1, [which], [0.0] * dim, [1.0] * dim, 1e-2, NAN, 0, cap, want_regions=False)
        val, err = r[0][0], r[1][0]
        if st.converged:
            assert abs(err / val) < 1e-2                      # the reference's only assertion, TestCubature.java:345
        if not st.ambiguous_decisions and which not in (2,):   # 2: coincidental exact ties (indicator function)
            assert (st.num_eval, st.num_regions) == (o.num_eval, o.num_regions)
            assert val == pytest.approx(o.val[0], rel=1e-12)
        if which in (0, 3, 6):                                 # smooth: the true error is inside the tolerance too
            assert abs(val - exact_integral(which, [1.0] * dim)) < 1e-2 * abs(val)


SPHERE_SRC = r"""
__device__ void sphere(const double* x, double* f, const double* p) {
    double r2 = 0.0;
    for (int d = 0; d < CUB_DIM; ++d) r2 += x[d] * x[d];
    f[0] = r2 < p[0] ? 1.0 : 0.0;   // note: r2 < radius, not radius^2, exactly as TestCubature.java:241
}
"""


def test_three_dim_unit_sphere(cb):
    """testThreeDimUnitSphere (TestCubature.java:223-280, commented out there as 'fails currently'): a user integrand
    as CUDA C source, Error.L2, relTol 1e-2.  It 'fails' for a reason: the degree-7/5 error estimate of a discontinuous
    integrand in 3-D shrinks so slowly that relTol 1e-2 is not reached within millions of evaluations (the oracle behaves
    the same).  Here the run gets a budget; the estimate must bracket the true volume and shrink with the budget."""
    ig = cb.Integrand.from_source(SPHERE_SRC, "sphere", 3, 1, [RADIUS])
    true_volume = (4.0 / 3.0) * math.pi * RADIUS ** 1.5 / 8.0         # octant of a sphere of radius sqrt(RADIUS)
    prev_err = None
    for budget in (200_000, 3_000_000):
        r, st, _ = cb.Cubature.integrate_ex(ig, [0.0] * 3, [1.0] * 3, 1e-2, NAN, cb.CubatureError.L2, budget)
        assert budget <= st.num_eval < budget + 2 * 33 or st.converged
        assert abs(r[0][0] - true_volume) <= r[1][0]
        if st.converged:
            assert abs(r[1][0] / r[0][0]) < 1e-2                      # TestCubature.java:279
        assert prev_err is None or r[1][0] < prev_err
        prev_err = r[1][0]


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_vector_valued_loop(cb, oracle, dim):
    """testCubature (TestCubature.java:353-432, commented out there): fdim = 1..max integrands at once, INDIVIDUAL.
    With the reference's ANY semantics the loop stops as soon as one component is converged (which is why the test
    was disabled); with strict_all every component has to meet the tolerance and both of its assertions hold for the
    smooth components."""
    max_fdim = 8 if dim == 3 else 7
    for fdim in range(1, max_fdim + 1):
        which = list(range(fdim))
        ig = cb.Integrand.builtin(cb.Family.REFTEST, dim, fdim, which)
        args = ([0.0] * dim, [1.0] * dim, 1e-2, NAN, cb.CubatureError.INDIVIDUAL, 4_000_000)
        # parity with the restated reference (ANY semantics), bit for bit
        r, st, tr = cb.Cubature.integrate_ex(ig, *args, select=cb.Select.HEAP_EXACT, trace_batches=4096)
        o = oracle.integrate(20, dim, fdim, which, *args[:2], 1e-2, NAN, 0, 4_000_000, want_regions=False)
        assert np.array_equal(r[0].view(np.uint64), o.val.view(np.uint64)) and st.num_eval == o.num_eval and tr.batches == o.batches
        # the documented semantics
        r, st, _ = cb.Cubature.integrate_ex(ig, *args, strict_all=True)
        if st.converged:
            for i in range(fdim):
                assert abs(r[1][i] / r[0][i]) < 1e-2                               # TestCubature.java:428
                if which[i] in (0, 3, 6):
                    assert abs(r[0][i] - exact_integral(which[i], [1.0] * dim)) < 1e-2 * abs(r[0][i])  # :427
