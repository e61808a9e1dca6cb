"""First contact with the B200: FP64 peak, bit parity of integrand / rule / full loop against the
oracle, and a first rule-kernel throughput number.  Prints; asserts nothing (pytest does that)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cubature_b200 as cb
from oracle import oracle as O
import cases

NAN = float("nan")
print("fp64 peak TFLOP/s:", cb.fp64_peak_tflops())

# 1. integrand bits
rng = np.random.default_rng(1)
for name, fam in cases.GENZ.items():
    for dim in (2, 5, 6):
        p = cases.genz_params(name, dim)
        x = rng.uniform(0, 1, (dim, 4096))
        ig = cb.Integrand.builtin(fam, dim, 1, p)
        g = ig.eval_points(x)
        c = O.eval_points(fam, dim, 1, p, x)
        print("points", name, dim, "ulp", cases.ulp_diff(g, c))

# 2. rule parity on random boxes
for name, fam in cases.GENZ.items():
    for dim in (2, 3, 6):
        p = cases.genz_params(name, dim)
        c, h = cases.random_boxes(dim, 1000, 7)
        ig = cb.Integrand.builtin(fam, dim, 1, p)
        gv, ge, gs, st = ig.eval_regions(c, h)
        cv, ce, cs = O.eval_regions(fam, dim, 1, p, c, h)
        print("rule", name, dim, "val ulp", cases.ulp_diff(gv, cv), "err ulp", cases.ulp_diff(ge, ce), "split mism", int((gs != cs).sum()))

# 3. full loop, heap-exact and device, vs oracle
def run(fam, dim, fdim, p, xmin, xmax, rel, ab, norm=0, maxEval=0, tr=None):
    o = O.integrate(fam, dim, fdim, p, xmin, xmax, rel, ab, norm, maxEval, transform=tr)
    ig = cb.Integrand.builtin(fam, dim, fdim, p)
    for sel in (cb.Select.HEAP_EXACT, cb.Select.DEVICE):
        t = time.time()
        r, st, trc = cb.Cubature.integrate_ex(ig, xmin, xmax, rel, ab, norm, maxEval, transform=tr, select=sel,
                                              trace_batches=4096, trace_regions=max(o.num_regions, 1))
        dt = time.time() - t
        print("  ", sel.name, "val ulp", cases.ulp_diff(r[0], o.val), "err ulp", cases.ulp_diff(r[1], o.err), "numEval", st.num_eval, o.num_eval,
              "regions", st.num_regions, o.num_regions, "batches_equal", trc.batches == o.batches, "split", trc.split_hist(dim), o.split_hist(dim),
              "amb", st.ambiguous_decisions, "dev_ms %.3f rule_ms %.3f wall %.3f launches %d" % (st.device_ms, st.rule_ms, dt * 1e3, st.gpu_launches))

print("KAT gauss3d")
run(10, 3, 1, [0.5], [-2] * 3, [2] * 3, 1e-4, 0.0)
print("C1 gauss3d sigma=1 1e-6")
run(10, 3, 1, [1.0], [-2] * 3, [2] * 3, 1e-6, 0.0)
for name, fam in cases.GENZ.items():
    print("genz", name, "d=5 relTol 1e-5")
    run(fam, 5, 1, cases.genz_params(name, 5), [0] * 5, [1] * 5, 1e-5, NAN, maxEval=3000000)
print("reftest dim1")
for w in range(8):
    run(20, 1, 1, [w], [0], [1], 1e-6, NAN, maxEval=100000)
print("reftest dim2 fdim2 norms")
for norm in range(5):
    run(20, 2, 2, [0, 4], [0, 0], [1, 1], 1e-3, NAN, norm)
print("oscvec dim1 fdim64 L2")
run(30, 1, 64, [], [0], [1], 1e-8, NAN, 2)
print("cexp dim3 fdim8 paired semi-inf")
run(31, 3, 8, [0.5, 1.0, 2.0, 3.0], [0] * 3, [float("inf")] * 3, 1e-4, NAN, 1, maxEval=2000000, tr=[1, 1, 1])

# 4. throughput of the rule kernel
for name in ("corner_peak", "gaussian", "oscillatory"):
    dim = 6
    ig = cb.Integrand.builtin(cases.GENZ[name], dim, 1, cases.genz_params(name, dim))
    for n in (1 << 20, 1 << 23):
        ms, cs = cb.bench_rule(ig, [0] * dim, [1] * dim, n, 5)
        print("bench", name, "n", n, "ms %.3f" % ms, "regions/s %.3e" % (n / ms * 1e3), "fevals/s %.3e" % (n * 149 / ms * 1e3), "checksum", cs)
