"""Burst vs sustained FP64 throughput on the GPU (power cap): DFMA probe and the rule kernel alone."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import cubature_b200 as cb, cases
if os.environ.get("PROBE_PEAK"):
    for w in (50, 800):
        b, s = cb.fp64_peak_sustained_tflops(-1, w)
        print("dfma probe window %%5d ms: burst %%.2f TFLOP/s  sustained %%.2f TFLOP/s" %% (w, b, s), flush=True)
fam, dim = os.environ.get("PROBE_FAMILY", "corner_peak"), int(os.environ.get("PROBE_DIM", "6"))
ig = cb.Integrand.builtin(cases.GENZ[fam], dim, 1, cases.genz_params(fam, dim))
n = 1 << 23
out = []
for rep in (5, 400):
    ms, cs = cb.bench_rule(ig, [0] * dim, [1] * dim, n, rep)
    out.append("%%4d reps: %%.3e regions/s" %% (rep, n / ms * 1e3))
print(fam, dim, " | ".join(out), flush=True)
''' % (ROOT, ROOT)
first = bool(os.environ.get("PROBE_PEAK"))
for defs in sys.argv[1:] or [""]:
    env = dict(os.environ, CUBATURE_JIT_DEFINES=defs, CUBATURE_BENCH_MODE="bisect")
    if not first:
        env.pop("PROBE_PEAK", None)
    first = False
    print("=== %r" % defs, end="  ", flush=True)
    subprocess.call([sys.executable, "-c", CODE], env=env)
