import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cubature_b200 as cb, cases
name, dim = sys.argv[1], int(sys.argv[2])
rel, max_eval = float(sys.argv[3]), int(float(sys.argv[4]))
p = cases.genz_params(name, dim)
ig = cb.Integrand.builtin(cases.GENZ[name], dim, 1, p)
r, s, t = cb.Cubature.integrate_ex(ig, [0.0] * dim, [1.0] * dim, rel, float("nan"), cb.CubatureError.INDIVIDUAL, max_eval, trace_batches=4096)
print(name, dim, r[0][0], r[1][0], s.num_regions, s.iterations, t.batches[-6:])
