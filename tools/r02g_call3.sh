#!/bin/bash
# Round-2 (late) GPU call 3: issue-slot probe, A/B of the Gray block size and the region-level range guard on the bench
# workload and on other families, then the GPU suite, the bench line and an ncu --set full capture of the rule kernel.
set -u
O=gpurun_out/r02g
mkdir -p $O
export PYTHONUNBUFFERED=1
./tools/issue_probe > $O/issue_probe.log 2>&1; cat $O/issue_probe.log
run() {  # defs, waves
    echo "== defs='$1' waves=$2"
    CUBATURE_JIT_DEFINES="$1" CUBATURE_RULE_WAVES=$2 timeout 150 python tools/c5_once.py 1e11 3 2>&1 | grep "call 2\|rror"
}
{
    run "" 16
    run "-DCUB_GM_BOX_GUARD=0" 16
    run "-DCUB_GM_GRAY_LOW=3" 16
    run "-DCUB_GM_GRAY_LOW=3 -DCUB_GM_BOX_GUARD=0" 16
    run "-DCUB_GM_GRAY_LOW=5" 16
    run "-DCUB_GM_MIN_BLOCKS=5" 16
    run "-DCUB_GM_PREFETCH=0" 16
} > $O/variants3.log 2>&1
cat $O/variants3.log
cat > /tmp/fam.py <<'P'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import cubature_b200 as cb, cases
for name, dim in [("product_peak", 6), ("gaussian", 6), ("oscillatory", 6), ("corner_peak", 5), ("corner_peak", 4)]:
    ig = cb.Integrand.builtin(cases.GENZ[name], dim, 1, cases.genz_params(name, dim))
    n = 1 << 23
    ms, cs = cb.bench_rule(ig, [0] * dim, [1] * dim, n, 5)
    print("%-14s d=%d  %.3f ms  %.3e regions/s  checksum %.17g" % (name, dim, ms, n / ms * 1e3, cs), flush=True)
P
for v in "" "-DCUB_GM_GRAY_LOW=3 -DCUB_GM_BOX_GUARD=0"; do echo "=== families, defs='$v'"; CUBATURE_BENCH_MODE=bisect CUBATURE_JIT_DEFINES="$v" timeout 300 python /tmp/fam.py; done > $O/families3.log 2>&1
cat $O/families3.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest3.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest3.log
timeout 600 python bench.py > $O/bench_n1_v3.json 2> $O/bench_n1_v3.err; echo "bench rc=$?"; cut -c1-400 $O/bench_n1_v3.json
export POOL_CAP=7e7
ncu --set full --clock-control none --import-source on -k "regex:cub_gm_scalar" --launch-skip 25 --launch-count 2 -f -o $O/ncu_gm3 timeout 600 python tools/c5_once.py 2e10 1 > $O/ncu_gm3.log 2>&1
ncu -i $O/ncu_gm3.ncu-rep --page raw --csv > $O/ncu_gm_scalar_v3_raw.csv 2>/dev/null
ncu -i $O/ncu_gm3.ncu-rep --page source --csv > $O/ncu_gm_scalar_v3_source.csv 2>/dev/null
rm -f $O/ncu_gm3.ncu-rep
tail -2 $O/ncu_gm3.log | cut -c1-200
