import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cubature_b200 as cb, cases
for name in sys.argv[1:]:
    dim = 5
    p = cases.genz_params(name, dim)
    ig = cb.Integrand.builtin(cases.GENZ[name], dim, 1, p)
    for rep in range(2):
        print(name, "call", rep, flush=True)
        r, s, t = cb.Cubature.integrate_ex(ig, [0.0] * dim, [1.0] * dim, 1e-8, float("nan"), cb.CubatureError.INDIVIDUAL, int(2e10))
        print("   device_ms %.2f rule_ms %.2f iterations %d regions %d" % (s.device_ms, s.rule_ms, s.iterations, s.num_regions), flush=True)
