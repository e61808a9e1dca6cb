import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import cubature_b200 as cb, cases
fam, dim = "corner_peak", 6
ig = cb.Integrand.builtin(cases.GENZ[fam], dim, 1, cases.genz_params(fam, dim))
n = 1 << 22
ms, cs = cb.bench_rule(ig, [0] * dim, [1] * dim, n, 2)
print("%.3e regions/s" % (n / ms * 1e3))
