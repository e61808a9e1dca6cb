"""Cost of maintaining the exact running sums inside the rule kernel: cub_gm_scalar (fused bisect mode, 6-D corner peak)
with / without the accumulators, for several grid shapes.  Run on the GPU box."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import cubature_b200 as cb, cases
fam, dim = "corner_peak", 6
ig = cb.Integrand.builtin(cases.GENZ[fam], dim, 1, cases.genz_params(fam, dim))
n = 1 << 23
ms, cs = cb.bench_rule(ig, [0] * dim, [1] * dim, n, 6)
print("%%.3e regions/s (%%.2f ms per %%d regions)" %% (n / ms * 1e3, ms, n), flush=True)
''' % (ROOT, ROOT)
VARIANTS = [
    ("no acc, one wave", {}),
    ("no acc, CTA per tile", {"CUBATURE_RULE_WAVES": "0"}),
    ("no acc, 16 waves", {"CUBATURE_RULE_WAVES": "16"}),
    ("no acc, 64 waves", {"CUBATURE_RULE_WAVES": "64"}),
    ("acc, window at 0 (far cells)", {"CUBATURE_BENCH_ACC": "1", "CUBATURE_RULE_WAVES": "0"}),
    ("acc, window 850, CTA per tile", {"CUBATURE_BENCH_ACC": "1", "CUBATURE_BENCH_ACC_EBASE": "850", "CUBATURE_BENCH_ACC_VBASE": "850", "CUBATURE_RULE_WAVES": "0"}),
    ("acc, window 850, one wave", {"CUBATURE_BENCH_ACC": "1", "CUBATURE_BENCH_ACC_EBASE": "850", "CUBATURE_BENCH_ACC_VBASE": "850"}),
    ("acc, window 850, 64 waves", {"CUBATURE_BENCH_ACC": "1", "CUBATURE_BENCH_ACC_EBASE": "850", "CUBATURE_BENCH_ACC_VBASE": "850", "CUBATURE_RULE_WAVES": "64"}),
]
for name, extra in VARIANTS:
    env = dict(os.environ, CUBATURE_BENCH_MODE="bisect", **extra)
    print("%-40s" % name, end=" ", flush=True)
    subprocess.call([sys.executable, "-c", CODE], env=env)
