#!/bin/bash
# ncu --set full captures of the round-2 kernels (run on a GPU box through gpurun; writes raw CSV pages to gpurun_out/).
# Each target program has been run without ncu first (the tests / bench of the same snapshot).
set -u
OUT=gpurun_out
mkdir -p $OUT
cap() {  # name, kernel regex, skip, count, command...
    local name=$1 kre=$2 skip=$3 count=$4
    shift 4
    ncu --set full --clock-control none --import-source on -k "regex:$kre" --launch-skip $skip --launch-count $count -f -o $OUT/$name "$@" > $OUT/$name.log 2>&1
    ncu -i $OUT/$name.ncu-rep --page raw --csv > $OUT/${name}_raw.csv 2>/dev/null
    rm -f $OUT/$name.ncu-rep
    tail -2 $OUT/$name.log | cut -c1-200
}
# scalar loop, 6-D corner peak, maxEval 2e10 (27 iterations, 6.7e7 regions): the largest decide / select launches
export POOL_CAP=7e7 CALLS=1  # pools allocated up front: no growth pauses, the launch sequence is fixed
cap r02_ncu_iter_large 'k_exact' 46 8 python tools/c5_once.py 2e10 1
# the rule kernel on the largest batch of the same job
cap r02_ncu_gm_scalar 'cub_gm_scalar' 25 2 python tools/c5_once.py 2e10 1
export POOL_CAP=4e5
# vector loop, BASELINE config C4: late iterations of the iteration kernel and of the staged rule kernel
cap r02_ncu_vec_iter 'k_vec_iter' 26 3 python tools/vec_timing.py C4
cap r02_ncu_rule_staged 'cub_rule_staged' 28 2 python tools/vec_timing.py C4
ls -la $OUT | tail -12
