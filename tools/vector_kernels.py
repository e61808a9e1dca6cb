"""Runs BASELINE configs C2 (GK15, fdim 1024) and C4 (d = 4, fdim 32, PAIRED, semi-infinite) once each: target for ncu captures
of the vector rule kernels (cub_gk_separable, cub_rule_staged)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cubature_b200 as cb  # noqa: E402

NAN, INF = float("nan"), float("inf")
which = sys.argv[1] if len(sys.argv) > 1 else "c4"
if which == "c2":
    ig = cb.Integrand.builtin(cb.Family.OSC_VEC, 1, 1024, [])
    r, st, _ = cb.Cubature.integrate_ex(ig, [0], [1], 1e-10, NAN, cb.CubatureError.L2, 0)
else:
    ig = cb.Integrand.builtin(cb.Family.CEXP, 4, 32, [0.25 * (m + 1) for m in range(16)])
    r, st, _ = cb.Cubature.integrate_ex(ig, [0] * 4, [INF] * 4, 1e-8, NAN, cb.CubatureError.PAIRED, 0, transform=[1] * 4)
print(which, "regions", st.num_regions, "iterations", st.iterations, "numEval", st.num_eval, "rule ms %.2f device ms %.2f" % (st.rule_ms, st.device_ms))
