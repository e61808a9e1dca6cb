"""Runs one scalar Genz case through the C ABI a few times and prints the statistics of the last call (for ncu launch
lists and quick timings on the GPU box):  python tools/run_case.py FAMILY DIM RELTOL MAXEVAL [CALLS]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cubature_b200 as cb  # noqa: E402
import cases  # noqa: E402

name, dim, rel, max_eval = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), int(float(sys.argv[4]))
calls = int(sys.argv[5]) if len(sys.argv) > 5 else 2
p = cases.genz_params(name, dim)
ig = cb.Integrand.builtin(cases.GENZ[name], dim, 1, p)
for c in range(calls):
    t0 = time.perf_counter()
    r, st, _ = cb.Cubature.integrate_ex(ig, [0.0] * dim, [1.0] * dim, rel, float("nan"), cb.CubatureError.INDIVIDUAL, max_eval)
    wall = time.perf_counter() - t0
    print("call %d: val %.17g err %.3e converged %d regions %d numEval %d iterations %d launches %d device %.3f ms rule %.3f ms wall %.3f ms ambiguous %d" % (
        c, r[0][0], r[1][0], st.converged, st.num_regions, st.num_eval, st.iterations, st.gpu_launches, st.device_ms, st.rule_ms, wall * 1e3,
        st.ambiguous_decisions), flush=True)
