#!/bin/bash
# Round-2 (late) GPU call 5, final tree: CTAs-per-SM / Gray-block sweep over the Genz families (information), the GPU
# suite, the bench line, and the ncu launch list of a one-step bench run.
set -u
O=gpurun_out/r02g
mkdir -p $O
export PYTHONUNBUFFERED=1
cat > /tmp/fam.py <<'P'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import cubature_b200 as cb, cases
for name in ["oscillatory", "product_peak", "corner_peak", "gaussian", "c0", "discontinuous"]:
    for dim in (5, 6):
        ig = cb.Integrand.builtin(cases.GENZ[name], dim, 1, cases.genz_params(name, dim))
        n = 1 << 22
        ms, cs = cb.bench_rule(ig, [0] * dim, [1] * dim, n, 5)
        print("%-14s d=%d  %.3f ms  %.3e regions/s  checksum %.17g" % (name, dim, ms, n / ms * 1e3, cs), flush=True)
P
for v in "" "-DCUB_GM_MIN_BLOCKS=4" "-DCUB_GM_MIN_BLOCKS=5" "-DCUB_GM_GRAY_LOW=5"; do echo "=== families, defs='$v'"; CUBATURE_BENCH_MODE=bisect CUBATURE_JIT_DEFINES="$v" timeout 300 python /tmp/fam.py; done > $O/families5.log 2>&1
cat $O/families5.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest5.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest5.log
timeout 600 python bench.py > $O/bench_n1_v5.json 2> $O/bench_n1_v5.err; echo "bench rc=$?"; cut -c1-300 $O/bench_n1_v5.json
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-time-to-tol --no-parity-check > $O/bench_short.json 2>&1; echo "short bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_bench_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-time-to-tol --no-parity-check > $O/ncu_launches.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py $O/launches_bench_final.csv | head -12
