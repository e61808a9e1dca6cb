"""BASELINE configs C2 and C4 (vector integrands): wall / device / rule time of repeated calls; CUBATURE_TIMING=1 adds the
phase time stamps of the iteration kernel."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cubature_b200 as cb
NAN, INF = float("nan"), float("inf")
E = cb.CubatureError
w16 = [0.25 * (m + 1) for m in range(16)]
CASES = {
    "C2": (cb.Family.OSC_VEC, 1, 1024, [], [0], [1], 1e-8, NAN, E.L2, None),
    "C4": (cb.Family.CEXP, 4, 32, w16, [0] * 4, [INF] * 4, 1e-8, NAN, E.PAIRED, [1] * 4),
}
for name in (sys.argv[1:] or ["C2", "C4"]):
    fam, dim, fdim, p, lo, hi, rel, ab, norm, tr = CASES[name]
    ig = cb.Integrand.builtin(fam, dim, fdim, p)
    for rep in range(int(os.environ.get("CALLS", "3"))):
        t0 = time.perf_counter()
        r, st, _ = cb.Cubature.integrate_ex(ig, lo, hi, rel, ab, norm, 0, transform=tr, pool_capacity=int(float(os.environ.get("POOL_CAP", "0"))))
        print("%s call %d: wall %.2f ms device %.2f ms rule %.2f ms iterations %d regions %d launches %d ambiguous %d" % (
            name, rep, (time.perf_counter() - t0) * 1e3, st.device_ms, st.rule_ms, st.iterations, st.num_regions, st.gpu_launches, st.ambiguous_decisions), flush=True)
