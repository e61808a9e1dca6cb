import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cubature_b200 as cb, cases
name = sys.argv[1] if len(sys.argv) > 1 else "corner_peak"
dim = 5
p = cases.genz_params(name, dim)
ig = cb.Integrand.builtin(cases.GENZ[name], dim, 1, p)
args = ([0.0] * dim, [1.0] * dim, 1e-8, float("nan"), cb.CubatureError.INDIVIDUAL, int(2e10))
for i in range(3):
    t0 = time.perf_counter()
    r, st, _ = cb.Cubature.integrate_ex(ig, *args)
    print("call", i, "wall ms %.3f device ms %.3f rule ms %.3f iters %d launches %d" % ((time.perf_counter() - t0) * 1e3, st.device_ms, st.rule_ms, st.iterations, st.gpu_launches), flush=True)
