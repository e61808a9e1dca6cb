#!/bin/bash
# Round-2 (late) GPU call: A/B of the rule kernel's geometry pipeline (CUB_GM_PREFETCH) and of the persistent-grid depth,
# then the GPU test suite, the bench line, a launch list and one ncu --set full capture of the rule kernel.
set -u
O=gpurun_out/r02g
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
run() {  # defs, waves
    echo "== defs='$1' waves=$2"
    CUBATURE_JIT_DEFINES="$1" CUBATURE_RULE_WAVES=$2 timeout 150 python tools/c5_once.py 1e11 3
}
{
    for w in 16 4 2 1; do run "-DCUB_GM_PREFETCH=0" $w; done
    for w in 16 4 2 1; do run "" $w; done
    for w in 16 2; do run "-DCUB_GM_MIN_BLOCKS=5" $w; done
    for w in 16 2; do run "-DCUB_GM_THREADS=64" $w; done
    for w in 16 2; do run "-DCUB_GM_THREADS=256 -DCUB_GM_MIN_BLOCKS=2" $w; done
} > $O/variants.log 2>&1
grep "==\|call 2" $O/variants.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; cut -c1-1500 $O/bench_n1.json
# rule kernel: ncu full capture of the two largest launches of a maxEval 2e10 job (run without ncu just above)
export POOL_CAP=7e7
ncu --set full --clock-control none --import-source on -k "regex:cub_gm_scalar" --launch-skip 25 --launch-count 2 -f -o $O/ncu_gm timeout 600 python tools/c5_once.py 2e10 1 > $O/ncu_gm.log 2>&1
ncu -i $O/ncu_gm.ncu-rep --page raw --csv > $O/ncu_gm_scalar_prefetch_raw.csv 2>/dev/null
ncu -i $O/ncu_gm.ncu-rep --page source --csv > $O/ncu_gm_scalar_prefetch_source.csv 2>/dev/null
rm -f $O/ncu_gm.ncu-rep
tail -2 $O/ncu_gm.log | cut -c1-200
ls -la $O
