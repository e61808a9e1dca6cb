#!/usr/bin/env python3
"""bench.py -- the headline metric of BASELINE.json on the hot path: Genz-Malik regions/s (and
f-evals/s) of the h-adaptive loop on the 6-D corner-peak Genz integrand (config C5).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference] [--max-eval E]

A "step" is one complete adaptive integration (`cubature_cuda_integrate` through the C ABI) of
    f(x) = (1 + sum a_i x_i)^-7  over [0,1]^6,  relTol 1e-13 (unreachable), maxEval E
so maxEval is the stop, exactly as BASELINE config C5 describes; E = 1e11 (the config's own value) is
the TOTAL over all GPUs at every N (strong scaling; the pool is sharded over the ranks).
  value   = regions evaluated by all ranks / device time (CUDA events on the library's stream, max over
            ranks); regions evaluated = numEval / 149.
  e2e     = the same through the public host API with host buffers (xmin/xmax/params in, [2][fdim] out),
            host<->device copies inside the timed region, host wall clock.
  roofline= FP64 pipe: algorithmic flops of the rule kernel (3 643 per 6-D corner-peak
            region) / its CUDA-event time, against the FP64 FMA peak measured live by a DFMA-chain
            microbenchmark (MEASURED_PEAKS.json has no FP64 entry); the HBM view is reported beside it.
  cpu_baseline = the CPU restatement of the reference (oracle/, kind "port": the Java reference cannot
            run here, there is no JVM) on a bounded sample of the same workload: one thread like the reference, and with
            the integrand batches on all host threads (the reference's vectorized interface); the faster one is reported.
`--impl reference` times that CPU restatement alone (rank 0 only).
Each step integrates a fresh pool far larger than the 126 MB L2 (3.4e8 regions x 124 B = 42 GB at E = 1e11),
so no L2 flush is needed between steps.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DIM = 6
P = 149
F_REGION = 3643.0      # algorithmic flops per evaluated 6-D corner-peak region
BYTES_REGION = 124.0   # algorithmic HBM bytes per evaluated region
TRAFFIC_REGION = 185.0 # DRAM bytes per evaluated region of the bisecting rule launch, ncu --set full (profiles/r02/late/ncu_gm_scalar_final_raw.csv: 3.44 GB read + 7.65 GB written for 5.99e7 regions)
NAN = float("nan")


def workload_params():
    import cases
    return cases.genz_params("corner_peak", DIM)


class ClockSampler:
    """Samples SM clocks / throttle reasons during the timed region (the quantities of the B200_PROFILING.md recipe's
    `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*` line) through NVML inside this process, every
    100 ms.  A polling `nvidia-smi -lms` child process perturbed the measurement itself: the same step took 130 ms
    without and 140 ... 165 ms with it (profiles/r02_summary.md); nvidia-smi is the fallback when pynvml is missing."""

    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.index = index
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop_flag = threading.Event()
        self.thread = None
        self.proc = None
        self.source = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.nv = pynvml
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.source = None
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self._physical_index()), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "250"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll_nvml(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for name, bit in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(float(os.environ.get("BENCH_SAMPLE_S", "0.1")))

    def _read_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.sm.append(float(r[0]))
                self.mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except (ValueError, IndexError):
                pass

    def stop(self):
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.thread:
            self.thread.join(timeout=2)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}


def cpu_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_port(max_eval, threads=1):
    """The oracle (CPU restatement of the reference) on the same integrand with a bounded maxEval.  threads > 1: the
    integrand batch of every iteration is evaluated by that many threads -- what the reference's README (129-163) invites the
    user of its vectorized interface to do; point generation, rule sums and the heap stay serial as in the reference."""
    from oracle import oracle as O
    p = workload_params()
    O.set_threads(threads)
    t0 = time.perf_counter()
    r = O.integrate(3, DIM, 1, p, [0.0] * DIM, [1.0] * DIM, 1e-13, NAN, 0, int(max_eval), want_regions=False)
    dt = time.perf_counter() - t0
    O.set_threads(1)
    return r.num_eval / P / dt, r.num_eval, dt


def cpu_baseline_record(sample):
    """Both flavours on the bounded sample; the faster one is the reported value."""
    t_all = cpu_threads()
    v1, _, dt1 = cpu_port(sample, 1)
    vt, _, dtt = cpu_port(sample, t_all) if t_all > 1 else (v1, 0, dt1)
    best, cores = (vt, t_all) if vt > v1 else (v1, 1)
    return {"value": best, "unit": "regions/s", "cores": cores, "kind": "port",
            "sample": "same integrand and loop with maxEval=%d; C++ restatement of the Java reference (no JVM in the image). "
                      "1 thread like the reference: %.4g regions/s (%.1f s); integrand batches on %d threads (the reference's "
                      "vectorized interface parallelised as its README suggests, everything else serial): %.4g regions/s (%.1f s)"
                      % (int(sample), v1, dt1, t_all, vt, dtt)}, dt1 + (dtt if t_all > 1 else 0.0)


def run_reference(args, rank):
    if rank != 0:
        return
    sample = int(args.cpu_sample)
    vals, times = [], []
    rec = None
    for i in range(args.warmup + args.steps):
        rec, dt = cpu_baseline_record(sample)
        if i >= args.warmup:
            vals.append(rec["value"])
            times.append(dt)
    value = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": "genz_malik_regions_per_s", "value": value, "unit": "regions/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(times), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "genz_corner_peak_d6 [0,1]^6 relTol=1e-13 maxEval=%d (bounded sample of the maxEval=1e11 job)" % sample},
        "fevals_per_s": value * P,
        "cpu_baseline": dict(rec, value=value),
        "e2e": {"value": value, "unit": "regions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def time_to_tol_block(cb, cases, dist, local_rank, comm, world, shard_min=0):
    """BASELINE's second headline: wall time to relTol 1e-8 for the six Genz families at dim 5 (config C3), at this number of
    GPUs (a pool is only sharded once it is large enough, so small problems run replicated).  Outside the timed region of the
    regions/s metric.  Per family: second call (the first one compiles the rule kernel), device time = max over ranks."""
    import torch
    out = {}
    dim = 5
    for name, fam in cases.GENZ.items():
        p = cases.genz_params(name, dim)
        ig = cb.Integrand.builtin(fam, dim, 1, p)
        args = ([0.0] * dim, [1.0] * dim, 1e-8, NAN, cb.CubatureError.INDIVIDUAL, 0)
        cb.Cubature.integrate_ex(ig, *args, device=local_rank, comm=comm, shard_min_regions=shard_min)
        best = None
        for _ in range(3):
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r, st, _ = cb.Cubature.integrate_ex(ig, *args, device=local_rank, comm=comm, shard_min_regions=shard_min)
            wall = (time.perf_counter() - t0) * 1e3
            t = torch.tensor([st.device_ms, wall, st.rule_ms], dtype=torch.float64, device="cuda")
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dev, wall, rule = [float(v) for v in t.cpu()]
            if best is None or wall < best["wall_ms"]:
                best = {"wall_ms": round(wall, 3), "device_ms": round(dev, 3), "rule_kernel_ms": round(rule, 3), "converged": bool(st.converged),
                        "regions": int(st.num_regions), "num_eval": int(st.num_eval), "iterations": int(st.iterations),
                        "val": float(r[0][0]), "err": float(r[1][0]), "ambiguous_decisions": int(st.ambiguous_decisions)}
        out[name] = best
    return out


def parity_block(cb, cases, dist, local_rank, comm, world, rank):
    """Every bench record doubles as a parity witness (outside the timed region).
    N > 1: the bench integrand with maxEval 3e9 on the sharded pool against the same call on this rank's GPU alone -- scalar
    integrands run on exact integer sums, so batches, numEval, value and error must be bit-identical.
    Every N: the device loop against the host-heap mode (the literal replay of the reference's loop with the JDK heap and its
    drifting sums) on a 5-D Gaussian -- same batches and counts, value to 1e-12."""
    import torch
    out = {}
    p = cases.genz_params("gaussian", 5)
    ig = cb.Integrand.builtin(cases.GENZ["gaussian"], 5, 1, p)
    args = ([0.0] * 5, [1.0] * 5, 1e-6, NAN, cb.CubatureError.INDIVIDUAL, 0)
    r1, s1, t1 = cb.Cubature.integrate_ex(ig, *args, device=local_rank, trace_batches=4096)
    r2, s2, t2 = cb.Cubature.integrate_ex(ig, *args, device=local_rank, trace_batches=4096, select=cb.Select.HEAP_EXACT)
    out["device_vs_host_heap"] = {"case": "genz_gaussian d=5 relTol=1e-6", "batches_equal": t1.batches == t2.batches,
                                  "numEval_equal": s1.num_eval == s2.num_eval, "regions_equal": s1.num_regions == s2.num_regions,
                                  "rel_diff": abs(r1[0][0] - r2[0][0]) / abs(r2[0][0]), "ambiguous_decisions": int(s1.ambiguous_decisions)}
    if world > 1:
        ig = cb.Integrand.builtin(cb.Family.GENZ_CORNER_PEAK, DIM, 1, workload_params())
        args = ([0.0] * DIM, [1.0] * DIM, 1e-13, NAN, cb.CubatureError.INDIVIDUAL, int(3e9))
        r1, s1, t1 = cb.Cubature.integrate_ex(ig, *args, device=local_rank, trace_batches=4096)
        dist.barrier()
        r2, s2, t2 = cb.Cubature.integrate_ex(ig, *args, device=local_rank, comm=comm, trace_batches=4096)
        ok = [t1.batches == t2.batches, s1.num_eval == s2.num_eval, r1[0][0] == r2[0][0] and r1[1][0] == r2[1][0]]
        flags = torch.tensor([0 if v else 1 for v in ok], device="cuda")
        dist.all_reduce(flags)
        flags = [int(v) == 0 for v in flags.cpu()]
        out["sharded_vs_one_gpu"] = {"case": "bench integrand, maxEval=3e9, %d ranks vs 1" % world, "batches_equal": flags[0], "numEval_equal": flags[1],
                                     "bit_identical": flags[2], "rel_diff": abs(r1[0][0] - r2[0][0]) / abs(r1[0][0]),
                                     "ambiguous_decisions": int(s2.ambiguous_decisions)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--max-eval", type=float, default=1e11, help="TOTAL maxEval of one step (all GPUs)")
    ap.add_argument("--cpu-sample", type=float, default=1.5e8, help="maxEval of the bounded CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-time-to-tol", action="store_true", help="skip the time-to-relTol-1e-8 block (six Genz families, dim 5)")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--shard-min", type=float, default=0, help="regions at which a pool is dealt out to the ranks (0 = library default)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "cuda":
        args.warmup = max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import cubature_b200 as cb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libcubature_cuda has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # ship the NCCL unique id of the library's own communicator from rank 0
        id_t = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            id_t = torch.frombuffer(bytearray(cb.Comm.unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(id_t, 0)
        comm = cb.Comm.init(bytes(id_t.cpu().numpy().tobytes()), world, rank, local_rank)

    params = workload_params()
    ig = cb.Integrand.builtin(cb.Family.GENZ_CORNER_PEAK, DIM, 1, params)
    max_eval = int(args.max_eval)
    xmin, xmax = [0.0] * DIM, [1.0] * DIM

    def step():
        t0 = time.perf_counter()
        r, st, _ = cb.Cubature.integrate_ex(ig, xmin, xmax, 1e-13, NAN, cb.CubatureError.INDIVIDUAL, max_eval, device=local_rank, comm=comm,
                                            shard_min_regions=int(args.shard_min))
        return r, st, time.perf_counter() - t0

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = cb.fp64_peak_tflops(local_rank)
    # the same probe back to back for 300 ms: on these boards the pure-DFMA probe does not reach the power cap, so
    # burst and sustained agree; a rule kernel that spills to local memory does (profiles/r01_summary.md, "power")
    fp64_burst, fp64_sustained = cb.fp64_peak_sustained_tflops(local_rank, 300.0) if rank == 0 else (fp64_peak, fp64_peak)
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("BENCH_NO_SAMPLER"):
        sampler.start()
    barrier()
    dev_ms, wall_s, rule_ms, launches, regions, h2d, d2h = [], [], [], 0, 0, 0, 0
    result = None
    for _ in range(args.steps):
        barrier()
        r, st, wall = step()
        barrier()
        result = r
        dev_ms.append(st.device_ms)
        wall_s.append(wall)
        rule_ms.append(st.rule_ms)
        launches += st.gpu_launches
        regions = st.num_eval // P          # all ranks (global numEval)
        iterations_last = st.iterations
        local_regions = st.regions_evaluated
        h2d += st.h2d_bytes
        d2h += st.d2h_bytes
    clocks = sampler.stop() if rank == 0 else None
    # one more (untimed) step with the batch log: algorithmic HBM bytes of the iteration kernels (decide / select)
    _, st_log, tr_log = cb.Cubature.integrate_ex(ig, xmin, xmax, 1e-13, NAN, cb.CubatureError.INDIVIDUAL, max_eval, device=local_rank, comm=comm,
                                                 shard_min_regions=int(args.shard_min), trace_batches=4096)
    iter_bytes, n_cut, n_now = 0.0, 0, 1
    for nr in tr_log.batches:
        k = nr // 2
        if k >= n_now:   # the whole pool is popped: the decide kernel adds 2 K children (16 B each), nothing else streams
            iter_bytes += 16.0 * 2 * k
        else:            # a batch is cut: + three passes over errmax / err, compaction (err, list, parent copies), parents subtracted
            iter_bytes += 16.0 * 2 * k + 3 * 8.0 * n_now + (8.0 * n_now + 4.0 * k + 32.0 * k) + 16.0 * k
            n_cut += 1
        n_now += k
    ambiguous_last = st.ambiguous_decisions
    import cases
    ttt = None if args.no_time_to_tol else time_to_tol_block(cb, cases, dist, local_rank, comm, world, int(args.shard_min))
    parity = None if args.no_parity_check else parity_block(cb, cases, dist, local_rank, comm, world, rank)

    t_dev = torch.tensor([sum(dev_ms), sum(wall_s) * 1e3, sum(rule_ms)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        ln = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(ln)
        launches = int(ln.item())
        per_rank = [torch.zeros(3, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(per_rank, torch.tensor([local_regions, sum(rule_ms) / len(rule_ms), sum(dev_ms) / len(dev_ms)], dtype=torch.float64, device="cuda"))
        per_rank = [[float(v) for v in t.cpu()] for t in per_rank]
    else:
        per_rank = [[float(local_regions), sum(rule_ms) / len(rule_ms), sum(dev_ms) / len(dev_ms)]]
    tot_dev_ms, tot_wall_ms, tot_rule_ms = [float(v) for v in t_dev.cpu()]
    if rank != 0:
        if comm is not None:
            comm.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    steps = args.steps
    value = regions * steps / (tot_dev_ms * 1e-3)
    e2e = regions * steps / (tot_wall_ms * 1e-3)
    # rule kernel (cub_gm_scalar): regions per launch-set = this rank's share; time = summed launch durations
    rule_regions_per_s = (regions / world) * steps / (tot_rule_ms * 1e-3)
    rule_launches_per_step = iterations_last + 1  # one fused bisect+rule launch per iteration, plus the root box
    achieved_tflops = rule_regions_per_s * F_REGION / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # nominal FP64 FMA peak: 64 FMA lanes per SM per clock at the maximum SM clock
    n_sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    fp64_nominal = n_sms * 64 * 2 * (clocks["sm_max_mhz"] or 1965.0) * 1e6 / 1e12 if clocks else None
    line = {
        "metric": "genz_malik_regions_per_s", "value": value, "unit": "regions/s", "n_gpus": world, "steps": steps,
        "warmup": args.warmup, "ms_per_step": tot_dev_ms / steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "genz_corner_peak_d6 [0,1]^6 relTol=1e-13 maxEval=%d total (BASELINE config C5), select=device" % max_eval,
                   "regions_per_step": regions, "points_per_region": P, "l2": "pool per step >> 126 MB L2, no flush needed",
                   "parallelism": ("pool sharded over %d GPUs; the decide / select kernels of every iteration exchange sums and histograms over "
                                   "NVLink peer memory (no NCCL call and no host decision in the loop)" % world) if world > 1 else "1 GPU",
                   "loop": "device-resident: decide / select / rule enqueued ahead in batches of 8 iterations, host reads a 512-byte control block "
                           "behind each batch while the next is queued",
                   "cpu_arm": "cpu_baseline and --impl reference time the CPU restatement of the reference on a bounded sample of this job "
                              "(maxEval=%d instead of %d): a stated baseline, not the target" % (int(args.cpu_sample), max_eval)},
        "fevals_per_s": value * P,
        "result": {"val": float(result[0][0]), "err": float(result[1][0])},
        "e2e": {"value": e2e, "unit": "regions/s", "h2d_bytes_per_step": h2d // steps, "d2h_bytes_per_step": d2h // steps,
                "ms_per_step": tot_wall_ms / steps},
        "gpu_launches": launches,
        "roofline": {"bound": "fp64", "kernel": "cub_gm_scalar", "achieved": achieved_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": achieved_tflops / fp64_peak,
                     # DRAM bytes per average rule launch: TRAFFIC_REGION per evaluated region measured with `ncu --set full`
                     # (dram__bytes_read.sum + dram__bytes_write.sum of one 6e7-region launch, profiles/r02/late/ncu_gm_scalar_final_raw.csv)
                     "traffic": TRAFFIC_REGION * regions / max(1, rule_launches_per_step),
                     "peak_source": "measured live: DFMA-chain microbenchmark (cubature_cuda_fp64_peak), best 10 ms launch; MEASURED_PEAKS.json has no FP64 entry",
                     "peak_sustained": fp64_sustained,
                     "peak_nominal": fp64_nominal, "frac_of_nominal": achieved_tflops / fp64_nominal if fp64_nominal else None,
                     "traffic_source": "constant: 185 B per evaluated region x regions per average launch, from the ncu --set full capture "
                                       "profiles/r02/late/ncu_gm_scalar_final_raw.csv (dram read + write of one bisecting launch: 124 B written per child + its parent's geometry); not measured in this run",
                     "flops_per_region": F_REGION, "rule_kernel_ms_per_step": tot_rule_ms / steps,
                     "rule_regions_per_s_per_gpu": rule_regions_per_s,
                     "hbm": {"achieved_gbs": rule_regions_per_s * BYTES_REGION / 1e9, "peak_gbs": hbm_peak,
                             "frac": rule_regions_per_s * BYTES_REGION / 1e9 / hbm_peak}},
        # the iteration kernels (k_exact_decide, k_exact_select): everything on the stream that is not a rule launch, launch gaps
        # included; bytes = what the algorithm has to stream (partition-wide), divided by the ranks
        "roofline_hbm": {"kernels": "k_exact_decide + k_exact_select", "non_rule_ms_per_step": (tot_dev_ms - tot_rule_ms) / steps,
                         "algorithmic_bytes_per_step": iter_bytes, "iterations": len(tr_log.batches), "iterations_that_cut": n_cut,
                         "achieved_gbs": iter_bytes / world / max(1e-9, (tot_dev_ms - tot_rule_ms) / steps * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                         "frac": iter_bytes / world / max(1e-9, (tot_dev_ms - tot_rule_ms) / steps * 1e-3) / 1e9 / hbm_peak,
                         "ncu": "profiles/r02/ncu_decide_select_large_raw.csv: decide 2.7 TB/s, select 2.4-2.6 TB/s of DRAM traffic on 3e7-region launches"},
        "clocks": clocks,
        # rank 0's device time of every timed step (a late host thread on one rank shows up as one long step on all ranks)
        "rank0_step_ms": [round(float(v), 3) for v in dev_ms],
        "per_rank": [{"regions_evaluated": int(p[0]), "rule_ms": p[1], "device_ms": p[2]} for p in per_rank],
        "ambiguous_decisions_per_step": int(ambiguous_last),
    }
    if ttt is not None:
        line["time_to_tol_ms"] = {"config": "BASELINE config C3: six Genz families, dim 5, relTol 1e-8, INDIVIDUAL, no maxEval; best wall of 3 calls after "
                                            "a compile call; device_ms = max over ranks", "n_gpus": world, "families": ttt}
    if parity is not None:
        line["parity_check"] = parity
    if not args.no_cpu_baseline:
        line["cpu_baseline"], _ = cpu_baseline_record(args.cpu_sample)
    print(json.dumps(line))
    if comm is not None:
        comm.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
