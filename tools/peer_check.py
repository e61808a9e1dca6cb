"""Multi-GPU check of the sharded device loop (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/peer_check.py
Every case is integrated on one GPU (each rank does that on its own device: identical bits) and on the pool sharded
over all ranks; scalar integrands (exact running sums): batches, numEval, region counts, value and error bit-identical;
vector integrands (host-driven sharded selection): identical batches and counts, values to 1e-12."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
import cubature_b200 as cb
import cases

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
id_t = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    id_t = torch.frombuffer(bytearray(cb.Comm.unique_id()), dtype=torch.uint8).cuda()
dist.broadcast(id_t, 0)
comm = cb.Comm.init(bytes(id_t.cpu().numpy().tobytes()), world, rank, local_rank)
NAN = float("nan")
CASES = [
    # name, dim, relTol, maxEval, shard_min
    ("corner_peak", 6, 1e-13, int(3e9), 0),
    ("corner_peak", 6, 1e-13, int(2e8), 4096),
    ("corner_peak", 5, 1e-8, 0, 0),
    ("product_peak", 5, 1e-8, 0, 1024),
    ("gaussian", 4, 1e-7, 0, 256),
    ("oscillatory", 3, 1e-9, 0, 64),
    ("discontinuous", 3, 1e-3, int(4e7), 64),
    ("c0", 4, 1e-5, int(2e8), 128),
]
# vector integrands (host-driven sharded selection over NCCL all-gathers): family id, dim, fdim, params, xmin, xmax, relTol, norm, maxEval
INF = float("inf")
VCASES = [
    (cb.Family.CEXP, 3, 8, [0.5, 1.0, 1.5, 2.0], [0.0] * 3, [INF] * 3, 1e-5, cb.CubatureError.PAIRED, 0),
    (cb.Family.OSC_VEC, 2, 5, [], [0.0] * 2, [1.0] * 2, 1e-8, cb.CubatureError.L1, 0),
]
bad = 0
if not os.environ.get("PEER_CASES"):
    for fam, dim, fdim, p, lo, hi, rel, norm, max_eval in VCASES:
        ig = cb.Integrand.builtin(fam, dim, fdim, p)
        tr = [1 if h == INF else 0 for h in hi] if INF in hi else None
        args = (lo, hi, rel, NAN, norm, max_eval)
        r1, s1, t1 = cb.Cubature.integrate_ex(ig, *args, transform=tr, device=local_rank, trace_batches=4096)
        dist.barrier()
        r2, s2, t2 = cb.Cubature.integrate_ex(ig, *args, transform=tr, device=local_rank, comm=comm, shard_min_regions=16, trace_batches=4096)
        same = (t1.batches == t2.batches and s1.num_eval == s2.num_eval and s1.num_regions == s2.num_regions and s1.converged == s2.converged)
        import numpy as np
        rel_v = float(np.max(np.abs(r1[0] - r2[0])) / max(float(np.max(np.abs(r1[0]))), 1e-300))
        ok = same and rel_v <= 1e-12
        flag = torch.tensor([0 if ok else 1], device="cuda")
        dist.all_reduce(flag)
        if rank == 0:
            print("vector family %d d=%d fdim=%d norm=%d: %s  regions=%d iters=%d conv=%d  |dv|/v=%.1e" % (
                int(fam), dim, fdim, int(norm), "OK" if flag.item() == 0 else "MISMATCH", s2.num_regions, s2.iterations, s2.converged, rel_v), flush=True)
        bad += int(flag.item())
if os.environ.get("PEER_CASES"):
    CASES = [CASES[int(i)] for i in os.environ["PEER_CASES"].split(",")]
for name, dim, rel, max_eval, smin in CASES:
    p = cases.genz_params(name, dim)
    ig = cb.Integrand.builtin(cases.GENZ[name], dim, 1, p)
    args = ([0.0] * dim, [1.0] * dim, rel, NAN, cb.CubatureError.INDIVIDUAL, max_eval)
    r1, s1, t1 = cb.Cubature.integrate_ex(ig, *args, device=local_rank, trace_batches=4096)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r2, s2, t2 = cb.Cubature.integrate_ex(ig, *args, device=local_rank, comm=comm, shard_min_regions=smin, trace_batches=4096)
    dt = time.perf_counter() - t0
    same = (t1.batches == t2.batches and s1.num_eval == s2.num_eval and s1.num_regions == s2.num_regions and s1.converged == s2.converged)
    rel_v = abs(r1[0][0] - r2[0][0]) / max(abs(r1[0][0]), 1e-300)
    rel_e = abs(r1[1][0] - r2[1][0]) / max(abs(r1[1][0]), 1e-300)
    # scalar integrands run on exact (integer) running sums: the decisions AND the reported sums are independent of how the
    # regions are spread over the ranks, so everything must be bit-identical -- no tolerance, no "ambiguous" allowance
    bits = r1[0][0] == r2[0][0] and r1[1][0] == r2[1][0]
    ok = same and bits and s1.iterations == s2.iterations
    amb = False
    flag = torch.tensor([0 if (ok or amb) else 1], device="cuda")
    dist.all_reduce(flag)
    if rank == 0:
        print("%-14s d=%d rel=%g maxEval=%d shard_min=%d: %s  regions=%d iters=%d conv=%d  |dv|/v=%.1e |de|/e=%.1e  single %.2f ms  sharded(%d) %.2f ms (wall %.2f)  ambiguous %d/%d"
              % (name, dim, rel, max_eval, smin, "OK" if flag.item() == 0 else "MISMATCH", s2.num_regions, s2.iterations, s2.converged, rel_v, rel_e,
                 s1.device_ms, world, s2.device_ms, dt * 1e3, s1.ambiguous_decisions, s2.ambiguous_decisions), flush=True)
        if not same:
            for i, (a, b) in enumerate(zip(t1.batches, t2.batches)):
                if a != b:
                    print("   first batch difference at iteration", i, a, b, "of", len(t1.batches), len(t2.batches))
                    break
    bad += int(flag.item())
# periodic rebalancing: pools dealt out very early, whose shards then grow unevenly; with a low trigger the
# loop pauses, moves the surplus regions and goes on -- results still bit-identical to one GPU
REBAL = [
    ("product_peak", 4, 1e-9, 0, 8),
    ("corner_peak", 3, 1e-11, int(4e8), 4),
    ("gaussian", 3, 1e-10, 0, 4),
    ("c0", 3, 1e-7, int(3e8), 4),
]
if not os.environ.get("PEER_CASES"):
    os.environ["CUBATURE_REBALANCE_MIN"] = "2000"
    total_rebalances = 0
    for name, dim, rel, max_eval, smin in REBAL:
        p = cases.genz_params(name, dim)
        ig = cb.Integrand.builtin(cases.GENZ[name], dim, 1, p)
        args = ([0.0] * dim, [1.0] * dim, rel, NAN, cb.CubatureError.INDIVIDUAL, max_eval)
        r1, s1, t1 = cb.Cubature.integrate_ex(ig, *args, device=local_rank, trace_batches=4096)
        dist.barrier()
        r2, s2, t2 = cb.Cubature.integrate_ex(ig, *args, device=local_rank, comm=comm, shard_min_regions=smin, trace_batches=4096)
        ok = (t1.batches == t2.batches and s1.num_eval == s2.num_eval and s1.num_regions == s2.num_regions and s1.converged == s2.converged
              and r1[0][0] == r2[0][0] and r1[1][0] == r2[1][0])
        flag = torch.tensor([0 if ok else 1], device="cuda")
        dist.all_reduce(flag)
        shares = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(shares, torch.tensor([s2.regions_evaluated], dtype=torch.int64, device="cuda"))
        if rank == 0:
            ev = [int(x.item()) for x in shares]
            print("rebalance %-14s d=%d rel=%g maxEval=%d shard_min=%d: %s  regions=%d iters=%d  rebalances=%d  regions evaluated per rank max/mean=%.3f"
                  % (name, dim, rel, max_eval, smin, "OK" if flag.item() == 0 else "MISMATCH", s2.num_regions, s2.iterations, s2.rebalances,
                     max(ev) * len(ev) / max(1, sum(ev))), flush=True)
        total_rebalances += s2.rebalances
        bad += int(flag.item())
    del os.environ["CUBATURE_REBALANCE_MIN"]
    if rank == 0:
        print("REBALANCES %d" % total_rebalances, flush=True)
comm.close()
dist.destroy_process_group()
sys.exit(1 if bad else 0)
