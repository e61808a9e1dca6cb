"""Rule-kernel tuning sweep on the GPU: regions/s of cub_gm_scalar for several JIT variants."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import cubature_b200 as cb, cases
for name, dim in [("corner_peak", 6), ("gaussian", 6), ("oscillatory", 6), ("product_peak", 6), ("corner_peak", 5), ("corner_peak", 3)]:
    ig = cb.Integrand.builtin(cases.GENZ[name], dim, 1, cases.genz_params(name, dim))
    n = 1 << 23
    ms, cs = cb.bench_rule(ig, [0] * dim, [1] * dim, n, 5)
    print("%%-14s d=%%d  %%.3f ms  %%.3e regions/s  %%.3e fevals/s" %% (name, dim, ms, n / ms * 1e3, n * cb.points_per_region(dim) / ms * 1e3), flush=True)
''' % (ROOT, ROOT)
for defs in sys.argv[1:] or [""]:
    env = dict(os.environ, CUBATURE_JIT_DEFINES=defs)
    print("=== CUBATURE_JIT_DEFINES=%r" % defs, flush=True)
    subprocess.call([sys.executable, "-c", CODE], env=env)
