"""Generates tests/golden/baseline_pins.json: the CPU oracle (oracle/cubature_oracle.cpp) at the settings BASELINE.json
states for its configs -- C3: the six Genz families, d = 5, relTol 1e-8; C4: d = 4, fdim = 32, PAIRED, semi-infinite,
relTol 1e-8; C5: 6-D corner peak, relTol 1e-13, truncated at maxEval 1e8 (the largest SURVEY.md 8d allows for parity).

The oracle needs minutes for these (1 to 70 s each with the integrand batches on 8 threads), so they are committed as
fixtures and the `-m gpu` tests compare against the file instead of running the oracle on the GPU box.  Every case is
run twice: as the reference computes (drifting incremental sums) and with rounding-free running sums
(oracle.set_exact_sums); `exact_sums_same` records that both take identical decisions, i.e. that no batch size of the
case is an artefact of the reference's accumulated rounding.  Run:  python tests/golden/make_baseline_golden.py
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import oracle as O  # noqa: E402
import cases  # noqa: E402

NAN = float("nan")
INF = float("inf")


def record(name, family, dim, fdim, params, xmin, xmax, rel, ab, norm=0, max_eval=0, transform=None):
    t0 = time.time()
    O.set_exact_sums(0)
    r = O.integrate(family, dim, fdim, params, xmin, xmax, rel, ab, norm, max_eval, transform=transform, want_regions=True)
    t1 = time.time()
    O.set_exact_sums(1)
    x = O.integrate(family, dim, fdim, params, xmin, xmax, rel, ab, norm, max_eval, transform=transform, want_regions=False)
    O.set_exact_sums(0)
    print("%-28s regions %9d numEval %11d iterations %3d  %.1f s  exact-sums variant: %s" % (
        name, r.num_regions, r.num_eval, r.iterations, t1 - t0, "same" if x.batches == r.batches else "DIFFERENT"), flush=True)
    return {
        "name": name, "family": family, "dim": dim, "fdim": fdim, "params": [float(p) for p in params],
        "xmin": [repr(float(v)) for v in xmin], "xmax": [repr(float(v)) for v in xmax],
        "relTol": None if rel != rel else rel, "absTol": None if ab != ab else ab, "norm": norm, "maxEval": max_eval,
        "transform": transform,
        "val_hex": [float(v).hex() for v in r.val], "err_hex": [float(v).hex() for v in r.err],
        "numEval": r.num_eval, "regions": r.num_regions, "iterations": r.iterations, "batches": r.batches,
        "split_hist": r.split_hist(dim), "converged": r.converged,
        "exact_sums_same": bool(x.batches == r.batches and x.num_eval == r.num_eval),
        "oracle_seconds": round(t1 - t0, 2),
    }


def main():
    O.set_threads(os.cpu_count() or 1)
    out = []
    for name, fam in cases.GENZ.items():
        out.append(record("c3_genz_%s_d5_rel1e-8" % name, fam, 5, 1, cases.genz_params(name, 5), [0] * 5, [1] * 5, 1e-8, NAN))
    out.append(record("c4_cexp_d4_f32_paired_rel1e-8", 31, 4, 32, [0.25 * (m + 1) for m in range(16)], [0] * 4, [INF] * 4, 1e-8,
                      NAN, 1, 0, transform=[1, 1, 1, 1]))
    out.append(record("c5_corner_d6_maxeval1e8", 3, 6, 1, cases.genz_params("corner_peak", 6), [0] * 6, [1] * 6, 1e-13, NAN, 0,
                      100000000))
    with open(os.path.join(HERE, "baseline_pins.json"), "w") as f:
        json.dump(out, f, indent=1, allow_nan=False)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
