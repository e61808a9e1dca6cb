#!/bin/bash
# Round-2 (late) GPU call 2: rule-kernel occupancy / CTA-size / grid-depth sweep on the bench workload (C5).
set -u
O=gpurun_out/r02g
mkdir -p $O
export PYTHONUNBUFFERED=1
run() {  # defs, waves, [threads override]
    echo "== defs='$1' waves=$2 threads=${3:-}"
    if [ -n "${3:-}" ]; then export CUBATURE_GM_THREADS=$3; else unset CUBATURE_GM_THREADS; fi
    CUBATURE_JIT_DEFINES="$1" CUBATURE_RULE_WAVES=$2 timeout 150 python tools/c5_once.py 1e11 3 2>&1 | grep "call 2\|rror"
}
{
    run "" 16
    run "" 32
    run "" 64
    run "" 0
    run "-DCUB_GM_THREADS=32 -DCUB_GM_MIN_BLOCKS=16" 16
    run "-DCUB_GM_THREADS=32 -DCUB_GM_MIN_BLOCKS=16" 64
    run "-DCUB_GM_THREADS=32 -DCUB_GM_MIN_BLOCKS=16 -DCUB_GM_PREFETCH=0" 16
    run "-DCUB_GM_THREADS=64 -DCUB_GM_MIN_BLOCKS=8" 16
    run "-DCUB_GM_THREADS=64 -DCUB_GM_MIN_BLOCKS=8" 64
    run "-DCUB_GM_GRAY_LOW=4" 16
    run "-DCUB_GM_GRAY_LOW=6" 16
    run "-DCUB_GM_THREADS=32 -DCUB_GM_MAXNREG=120" 16 32
    run "-DCUB_GM_THREADS=64 -DCUB_GM_MAXNREG=112" 16 64
    run "-DCUB_GM_THREADS=32 -DCUB_GM_MAXNREG=136" 16 32
    run "-DCUB_GM_THREADS=32 -DCUB_GM_MAXNREG=144" 16 32
    run "-DCUB_GM_THREADS=32 -DCUB_GM_MAXNREG=168" 16 32
} > $O/variants2.log 2>&1
cat $O/variants2.log
