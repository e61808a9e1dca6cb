"""Runs pins of tests/golden/baseline_pins.json through DEVICE mode and prints what differs (debugging aid)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cubature_b200 as cb
NAN = float("nan")
pins = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "baseline_pins.json")))
want = sys.argv[1:] or [c["name"] for c in pins]
for c in pins:
    if not any(w in c["name"] for w in want):
        continue
    rel = NAN if c["relTol"] is None else c["relTol"]
    ab = NAN if c["absTol"] is None else c["absTol"]
    ig = cb.Integrand.builtin(c["family"], c["dim"], c["fdim"], c["params"])
    for rep in range(2):
        t0 = time.time()
        try:
            r, st, tr = cb.Cubature.integrate_ex(ig, [float(v) for v in c["xmin"]], [float(v) for v in c["xmax"]], rel, ab, c["norm"], c["maxEval"],
                                                 transform=c["transform"], trace_batches=8192)
        except RuntimeError as e:
            print(c["name"], "FAILED", e)
            break
        dt = time.time() - t0
        ok = (st.num_eval, st.num_regions, st.iterations) == (c["numEval"], c["regions"], c["iterations"]) and tr.batches == c["batches"]
        print("%-36s %s regions %d/%d numEval %d/%d iters %d/%d amb %d launches %d device %.2f ms rule %.2f ms wall %.2f ms" % (
            c["name"], "OK " if ok else "DIFF", st.num_regions, c["regions"], st.num_eval, c["numEval"], st.iterations, c["iterations"],
            st.ambiguous_decisions, st.gpu_launches, st.device_ms, st.rule_ms, dt * 1e3), flush=True)
        if not ok:
            for i, (a, b) in enumerate(zip(tr.batches, c["batches"])):
                if a != b:
                    print("   first differing batch %d: %d vs %d" % (i, a, b))
                    break
