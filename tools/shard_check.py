import os, sys, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cubature_b200 as cb, cases
name, dim, rel, cap = "corner_peak", 6, 1e-13, int(3e9)
p = cases.genz_params(name, dim)
ig = cb.Integrand.builtin(cases.GENZ[name], dim, 1, p)
args = ([0.0] * dim, [1.0] * dim, rel, float("nan"), cb.CubatureError.INDIVIDUAL, cap)
r1, s1, t1 = cb.Cubature.integrate_ex(ig, *args, trace_batches=256)
print("single:", repr(r1[0][0]), repr(r1[1][0]), s1.num_regions, s1.ambiguous_decisions, t1.batches)
for R, smin in ((2, 0), (2, 4096), (4, 0)):
    out = [None] * R
    def work(r):
        comm = cb.Comm.init_local(777 + R * 10 + smin % 7, R, r)
        out[r] = cb.Cubature.integrate_ex(ig, *args, comm=comm, shard_min_regions=smin, trace_batches=256)
        comm.close()
    ts = [threading.Thread(target=work, args=(r,)) for r in range(R)]
    [t.start() for t in ts]; [t.join() for t in ts]
    rr, ss, tt = out[0]
    print("R=%d shard_min=%d:" % (R, smin), repr(rr[0][0]), repr(rr[1][0]), ss.num_regions, ss.ambiguous_decisions, "batches equal:", tt.batches == t1.batches)
    if tt.batches != t1.batches:
        for i, (a, b) in enumerate(zip(tt.batches, t1.batches)):
            if a != b:
                print("   first diff at iteration", i, a, b); break
