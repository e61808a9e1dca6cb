"""One C5-style call (6-D corner peak, relTol 1e-13) at a given maxEval after a warm-up call at a small one: the ncu target
for the iteration kernels (decide / select) on a large pool."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cubature_b200 as cb, cases
max_eval = int(float(sys.argv[1])) if len(sys.argv) > 1 else int(2e10)
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ig = cb.Integrand.builtin(cb.Family.GENZ_CORNER_PEAK, 6, 1, cases.genz_params("corner_peak", 6))
for i in range(calls):
    # POOL_CAP: allocate the pool up front (no growth pauses, so the launch sequence is exactly decide, select, rule per iteration)
    r, st, _ = cb.Cubature.integrate_ex(ig, [0.0] * 6, [1.0] * 6, 1e-13, float("nan"), cb.CubatureError.INDIVIDUAL, max_eval,
                                        pool_capacity=int(float(os.environ.get("POOL_CAP", "0"))))
    print("call %d: device %.2f ms rule %.2f ms iterations %d regions %d launches %d" % (i, st.device_ms, st.rule_ms, st.iterations, st.num_regions, st.gpu_launches), flush=True)
