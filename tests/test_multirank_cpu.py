"""CPU tests of the N > 1 path: world_size-2 (and 3) `gloo` process groups exercising the host-side
sharding arithmetic of the multi-GPU driver (cubature_cuda_shard_count / cubature_cuda_tie_share, the
functions integrate_device itself calls) and the rank plumbing bench.py uses (unique-id broadcast)."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import cubature_b200 as cb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = cb._ffi.lib()
    ok = True
    # 1. the id broadcast bench.py performs (128 opaque bytes from rank 0)
    ident = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        ident = torch.arange(128, dtype=torch.uint8)
    dist.broadcast(ident, 0)
    ok &= bool((ident == torch.arange(128, dtype=torch.uint8)).all())
    # 2. dealing a replicated pool out: the shards are disjoint, cover everything and are balanced
    rng = np.random.default_rng(5)
    for n in [0, 1, 2, world, world + 1, 4096 * world, 4096 * world + 3, 12345]:
        mine = L.cubature_cuda_shard_count(n, world, rank)
        t = torch.tensor([mine], dtype=torch.int64)
        dist.all_reduce(t)
        ok &= int(t.item()) == n
        slots = torch.zeros(n + 1, dtype=torch.int64)
        for j in range(mine):
            slots[j * world + rank] += 1
        dist.all_reduce(slots)
        ok &= bool((slots[:n] == 1).all())
    # 3. tied regions at the threshold are popped in (rank, slot) order
    for trial in range(200):
        ties = torch.tensor(rng.integers(0, 7, world), dtype=torch.int64)
        dist.broadcast(ties, 0)
        total = int(ties.sum())
        take = torch.tensor([int(rng.integers(0, total + 1))], dtype=torch.int64)
        dist.broadcast(take, 0)
        arr = (C.c_int64 * world)(*ties.tolist())
        share = L.cubature_cuda_tie_share(int(take.item()), arr, world, rank)
        ok &= 0 <= share <= int(ties[rank])
        shares = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(shares, torch.tensor([share], dtype=torch.int64))
        shares = [int(s.item()) for s in shares]
        ok &= sum(shares) == int(take.item())
        # lower ranks are served first: a rank takes nothing unless all lower ranks are exhausted
        for r in range(1, world):
            if shares[r] > 0:
                ok &= all(shares[q] == int(ties[q]) for q in range(r))
    # 4. bad arguments
    ok &= L.cubature_cuda_shard_count(10, world, world) < 0
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok))


@pytest.mark.parametrize("world", [2, 3])
def test_sharding_logic_under_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in results) == list(range(world))
    assert all(ok for _, ok in results)


def test_reference_arm_runs_on_rank0_only():
    """bench.py --impl reference under a 2-rank launch: rank 0 prints one JSON line, rank 1 prints nothing."""
    import json
    import subprocess
    outs = []
    for rank in (0, 1):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                            "--warmup", "0", "--cpu-sample", "200000"], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout.strip())
    line = json.loads(outs[0])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and outs[1] == ""
