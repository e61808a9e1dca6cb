"""The five BASELINE.json configurations on one B200 (C5 truncated to maxEval 3e9 here; bench.py runs it in full): second-call
wall time through the C ABI, counts, and a comparison with the CPU oracle (same counts, values to 1e-12) where the oracle
finishes in seconds.  Prints a markdown table."""
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import cubature_b200 as cb  # noqa: E402
import cases  # noqa: E402
from oracle import oracle as O  # noqa: E402

NAN, INF = float("nan"), float("inf")
E = cb.CubatureError
w16 = [0.25 * (m + 1) for m in range(16)]
CONFIGS = [
    # label, family, dim, fdim, params, xmin, xmax, relTol, absTol, norm, maxEval, transform (None: derived from infinite bounds), oracle?
    ("C1 test: exp(-0.5|x|^2) [-2,2]^3 relTol 1e-4 (reference KAT)", cb.Family.GAUSS_ISO, 3, 1, [0.5], [-2] * 3, [2] * 3, 1e-4, 0.0, E.INDIVIDUAL, 0, True),
    ("C1: exp(-|x|^2) [-2,2]^3 relTol 1e-6", cb.Family.GAUSS_ISO, 3, 1, [1.0], [-2] * 3, [2] * 3, 1e-6, 0.0, E.INDIVIDUAL, 0, True),
    ("C2: GK15 d=1 fdim=1024 cos((k+1)x) [0,1] L2 relTol 1e-8", cb.Family.OSC_VEC, 1, 1024, [], [0], [1], 1e-8, NAN, E.L2, 0, True),
    ("C4: d=4 fdim=32 exp(-(1+i w_m) sum x) [0,inf)^4 PAIRED relTol 1e-8", cb.Family.CEXP, 4, 32, w16, [0] * 4, [INF] * 4, 1e-8, NAN, E.PAIRED, 0, False),
    ("C4 (relTol 1e-5, oracle-checked)", cb.Family.CEXP, 4, 32, w16, [0] * 4, [INF] * 4, 1e-5, NAN, E.PAIRED, 0, True),
    ("C5 truncated: corner peak d=6 maxEval 3e9", cb.Family.GENZ_CORNER_PEAK, 6, 1, list(cases.genz_params("corner_peak", 6)), [0] * 6, [1] * 6, 1e-13, NAN, E.INDIVIDUAL, int(3e9), False),
]
print("| config | converged | numEval | regions | iterations | GPU wall ms (2nd call) | rule ms | max rel. diff vs oracle (val, err) | counts = oracle | CPU oracle s |")
print("|---|---|---|---|---|---|---|---|---|---|")
for label, fam, dim, fdim, p, lo, hi, rel, ab, norm, max_eval, check in CONFIGS:
    ig = cb.Integrand.builtin(fam, dim, fdim, p)
    tr = [1 if math.isinf(h) else 0 for h in hi] if any(math.isinf(h) for h in hi) else None  # 1 = semi-infinite upwards
    cb.Cubature.integrate_ex(ig, lo, hi, rel, ab, norm, max_eval, transform=tr)
    t0 = time.perf_counter()
    r, st, _ = cb.Cubature.integrate_ex(ig, lo, hi, rel, ab, norm, max_eval, transform=tr)
    wall = time.perf_counter() - t0
    diff, same, cpu = "", "", ""
    if check:
        t0 = time.perf_counter()
        o = O.integrate(int(fam), dim, fdim, p, lo, hi, rel, ab, int(norm), max_eval, transform=tr, want_regions=False)
        cpu = "%.2f" % (time.perf_counter() - t0)
        dv = float(np.max(np.abs(r[0] - o.val)) / max(np.max(np.abs(o.val)), 1e-300))
        de = float(np.max(np.abs(r[1] - o.err)) / max(np.max(np.abs(o.err)), 1e-300))
        diff = "%.1e, %.1e" % (dv, de)
        same = str(o.num_eval == st.num_eval and o.num_regions == st.num_regions)
    print("| %s | %s | %d | %d | %d | %.2f | %.2f | %s | %s | %s |" % (label, bool(st.converged), st.num_eval, st.num_regions, st.iterations,
                                                                     wall * 1e3, st.rule_ms, diff, same, cpu), flush=True)
