#!/bin/bash
# Round-2 (late) GPU call 7: region-level guard of the exponential families (dm_exp_inrange) A/B, then the GPU suite and the bench line.
set -u
O=gpurun_out/r02g
mkdir -p $O
export PYTHONUNBUFFERED=1
cat > /tmp/fam.py <<'P'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import cubature_b200 as cb, cases
for name in ["gaussian", "c0", "discontinuous", "oscillatory"]:
    for dim in (5, 6):
        ig = cb.Integrand.builtin(cases.GENZ[name], dim, 1, cases.genz_params(name, dim))
        n = 1 << 22
        ms, cs = cb.bench_rule(ig, [0] * dim, [1] * dim, n, 5)
        print("%-14s d=%d  %.3f ms  %.3e regions/s  checksum %.17g" % (name, dim, ms, n / ms * 1e3, cs), flush=True)
ig = cb.Integrand.builtin(cb.Family.GAUSS_ISO, 3, 1, [1.0])
ms, cs = cb.bench_rule(ig, [-2.0] * 3, [2.0] * 3, 1 << 22, 5)
print("%-14s d=%d  %.3f ms  %.3e regions/s  checksum %.17g" % ("gauss_iso", 3, ms, (1 << 22) / ms * 1e3, cs), flush=True)
P
for v in "" "-DCUB_GM_BOX_GUARD=0"; do echo "=== families, defs='$v'"; CUBATURE_BENCH_MODE=bisect CUBATURE_JIT_DEFINES="$v" timeout 300 python /tmp/fam.py; done > $O/families7.log 2>&1
cat $O/families7.log
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest7.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest7.log
timeout 300 python bench.py --no-cpu-baseline > $O/bench_n1_v7.json 2> $O/bench_n1_v7.err; echo "bench rc=$?"; cut -c1-260 $O/bench_n1_v7.json
