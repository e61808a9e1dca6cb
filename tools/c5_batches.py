"""Pool size, batch size and kept regions of every iteration of the bench job (6-D corner peak, relTol 1e-13)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cubature_b200 as cb, cases
max_eval = int(float(sys.argv[1])) if len(sys.argv) > 1 else int(1e11)
ig = cb.Integrand.builtin(cb.Family.GENZ_CORNER_PEAK, 6, 1, cases.genz_params("corner_peak", 6))
r, st, tr = cb.Cubature.integrate_ex(ig, [0.0] * 6, [1.0] * 6, 1e-13, float("nan"), cb.CubatureError.INDIVIDUAL, max_eval, trace_batches=4096)
n = 1
print("| iteration | pool | popped | kept | popped / pool |\n|---|---|---|---|---|")
for i, nr in enumerate(tr.batches):
    k = nr // 2
    print("| %d | %d | %d | %d | %.4f |" % (i + 1, n, k, n - k, k / n))
    n += k
