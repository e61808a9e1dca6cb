#!/bin/bash
# Round-2 (late) GPU call 6: the committed tree once more -- GPU suite, smoke, bench line.
set -u
O=gpurun_out/r02g
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest6.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest6.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > $O/bench_n1_v6.json 2> $O/bench_n1_v6.err; echo "bench rc=$?"; cut -c1-260 $O/bench_n1_v6.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_v6.json 2>&1; echo "ref rc=$?"; cut -c1-300 $O/bench_ref_v6.json
