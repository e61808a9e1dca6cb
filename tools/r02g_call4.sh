#!/bin/bash
# Round-2 (late) GPU call 4 (2 GPUs): the multi-GPU parity test the 1-GPU suite skips, and the 2-GPU bench line.
set -u
O=gpurun_out/r02g
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -x -q -k "sharded_pool_over_nvlink or peers" > $O/pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > $O/bench_n2_v3.json 2> $O/bench_n2_v3.err; echo "bench rc=$?"; cut -c1-300 $O/bench_n2_v3.json; grep -o '"parity_check": {[^}]*}[^}]*}' $O/bench_n2_v3.json | cut -c1-600
