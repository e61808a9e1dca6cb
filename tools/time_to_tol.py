"""BASELINE config C3: Genz test family, dim = 5, relTol 1e-8 on one B200 -- wall time to tolerance,
regions/s and f-evals/s per family, next to the CPU oracle (C++ restatement of the reference, 1 thread)
where that finishes in reasonable time.  Prints a markdown table."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cubature_b200 as cb  # noqa: E402
import cases  # noqa: E402
from oracle import oracle as O  # noqa: E402

NAN = float("nan")
dim = int(sys.argv[1]) if len(sys.argv) > 1 else 5
rel = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-8
cap = int(float(sys.argv[3])) if len(sys.argv) > 3 else int(2e10)
cpu_cap = int(float(sys.argv[4])) if len(sys.argv) > 4 else int(6e8)
P = cb.points_per_region(dim)
print("| family | converged | numEval | regions | iterations | GPU wall ms (2nd call) | device ms | rule ms | regions/s | f-evals/s | |val-exact|/exact | est. rel. err | CPU oracle s | speed-up |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for name, fam in cases.GENZ.items():
    p = cases.genz_params(name, dim)
    ig = cb.Integrand.builtin(fam, dim, 1, p)
    args = ([0.0] * dim, [1.0] * dim, rel, NAN, cb.CubatureError.INDIVIDUAL, cap)
    cb.Cubature.integrate_ex(ig, *args)  # first call: JIT + allocations
    t0 = time.perf_counter()
    r, st, _ = cb.Cubature.integrate_ex(ig, *args)
    wall = time.perf_counter() - t0
    exact = cases.genz_exact(name, dim, p)
    cpu = ""
    sp = ""
    if st.num_eval <= cpu_cap:
        t0 = time.perf_counter()
        o = O.integrate(fam, dim, 1, p, *args[:2], rel, NAN, 0, cap, want_regions=False)
        ct = time.perf_counter() - t0
        cpu = "%.2f" % ct
        sp = "%.0fx" % (ct / wall)
        assert o.num_eval == st.num_eval and o.num_regions == st.num_regions, (name, o.num_eval, st.num_eval)
    print("| %s | %s | %d | %d | %d | %.2f | %.2f | %.2f | %.3e | %.3e | %.1e | %.1e | %s | %s |" % (
        name, bool(st.converged), st.num_eval, st.num_regions, st.iterations, wall * 1e3, st.device_ms, st.rule_ms,
        st.num_eval / P / wall, st.num_eval / wall, abs(r[0][0] - exact) / abs(exact), r[1][0] / abs(r[0][0]) if r[0][0] else 0, cpu, sp), flush=True)
