#!/bin/bash
# ncu --set full of the vector loop's kernels on late iterations of BASELINE config C4 (pool allocated up front: no growth
# pauses, launch i of k_vec_iter / cub_rule_staged is iteration i)
set -u
OUT=gpurun_out
mkdir -p $OUT
export POOL_CAP=1e6 CALLS=1
for spec in "r02_ncu_vec_iter k_vec_iter 26 3" "r02_ncu_rule_staged cub_rule_staged 27 3"; do
    set -- $spec
    ncu --set full --clock-control none --import-source on -k "regex:$2" --launch-skip $3 --launch-count $4 -f -o $OUT/$1 python tools/vec_timing.py C4 > $OUT/$1.log 2>&1
    ncu -i $OUT/$1.ncu-rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null
    rm -f $OUT/$1.ncu-rep
    tail -2 $OUT/$1.log | cut -c1-200
done
