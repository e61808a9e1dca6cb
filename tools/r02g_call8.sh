#!/bin/bash
# Round-2 (late) GPU call 8: small batches on the point-parallel staged kernel -- GPU suite, then time to tolerance with and without.
set -u
O=gpurun_out/r02g
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 115 python -m pytest tests -m gpu -x -q > $O/pytest8.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest8.log
cat > /tmp/tt.py <<'P'
import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import cubature_b200 as cb, cases
NAN = float("nan")
for name in ["discontinuous", "oscillatory", "c0", "corner_peak", "product_peak", "gaussian"]:
    ig = cb.Integrand.builtin(cases.GENZ[name], 5, 1, cases.genz_params(name, 5))
    args = ([0.0] * 5, [1.0] * 5, 1e-8, NAN, cb.CubatureError.INDIVIDUAL, int(2e10))
    out = []
    for small in ("2048", "0", "512", "8192"):
        os.environ["CUBATURE_SMALL_BATCH"] = small
        best = None
        for rep in range(3):
            r, st, _ = cb.Cubature.integrate_ex(ig, *args)
            if rep and (best is None or st.device_ms < best[0]): best = (st.device_ms, st.rule_ms, st.num_eval, st.gpu_launches, r[0][0])
        out.append("small=%s device %.3f ms rule %.3f ms numEval %d launches %d val %.17g" % ((small,) + best))
    print(name, " | ".join(out), flush=True)
P
timeout 60 python /tmp/tt.py > $O/small_batch_time_to_tol.log 2>&1; cat $O/small_batch_time_to_tol.log | cut -c1-420
