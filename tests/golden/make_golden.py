"""Generates tests/golden/oracle_pins.json from the CPU oracle (oracle/cubature_oracle.cpp).

These are REGRESSION pins of the restatement (region counts, numEval, batch sizes, split-dimension
histograms, value/error bits), not outputs of the Java reference: the reference cannot run in this
image (no JVM).  The first block repeats the "derived pins" of SURVEY.md Appendix E, which an
independent throwaway restatement produced during the survey; agreement with them is checked by
tests/test_oracle.py.  Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import oracle as O  # noqa: E402
import cases  # noqa: E402

NAN = float("nan")


def record(name, family, dim, fdim, params, xmin, xmax, rel, ab, norm=0, max_eval=0, transform=None):
    r = O.integrate(family, dim, fdim, params, xmin, xmax, rel, ab, norm, max_eval, transform=transform)
    # exact ties between the error estimates of distinct final regions: the order in which the JDK heap
    # pops tied regions is not reproduced by the device-side selection, so tests relax there
    import numpy as np
    em = r.errmax[r.errmax > 0]
    ties = int(len(em) - len(np.unique(em)))
    return {
        "exact_ties": ties,
        "name": name, "family": family, "dim": dim, "fdim": fdim, "params": [float(p) for p in params],
        "xmin": [repr(float(v)) for v in xmin], "xmax": [repr(float(v)) for v in xmax],  # strings: +-inf allowed
        "relTol": None if rel != rel else rel, "absTol": None if ab != ab else ab, "norm": norm, "maxEval": max_eval,
        "transform": transform,
        "val_hex": [float(v).hex() for v in r.val], "err_hex": [float(v).hex() for v in r.err],
        "numEval": r.num_eval, "regions": r.num_regions, "iterations": r.iterations, "batches": r.batches,
        "split_hist": r.split_hist(dim), "converged": r.converged,
    }


def main():
    inf = float("inf")
    out = []
    out.append(record("kat_gauss3d_sigma0.5_rel1e-4", 10, 3, 1, [0.5], [-2] * 3, [2] * 3, 1e-4, 0.0))
    out.append(record("c1_gauss3d_sigma1_rel1e-6", 10, 3, 1, [1.0], [-2] * 3, [2] * 3, 1e-6, 0.0))
    out.append(record("gauss3d_sigma0.5_rel1e-6", 10, 3, 1, [0.5], [-2] * 3, [2] * 3, 1e-6, 0.0))
    for dim in (1, 2):
        for w in range(8):
            out.append(record("reftest_%d_dim%d" % (w, dim), 20, dim, 1, [w], [0] * dim, [1] * dim, 1e-2, NAN))
    for dim in (3, 4):
        for w in (0, 3, 4, 6, 8 if dim == 3 else 7):
            out.append(record("reftest_%d_dim%d" % (w, dim), 20, dim, 1, [w], [0] * dim, [1] * dim, 1e-3, NAN, max_eval=400000))
    for norm in range(5):
        out.append(record("norm%d_fdim2" % norm, 20, 2, 2, [0, 4], [0, 0], [1, 1], 1e-3, NAN, norm))
        out.append(record("norm%d_fdim3_abs" % norm, 20, 2, 3, [0, 4, 6], [0, 0], [1, 1], NAN, 1e-4, norm))
    for name, fam in cases.GENZ.items():
        for dim in (2, 3, 5):
            out.append(record("genz_%s_d%d" % (name, dim), fam, dim, 1, cases.genz_params(name, dim), [0] * dim, [1] * dim,
                              1e-5, NAN, max_eval=2000000))
    out.append(record("genz_corner_d6_maxeval", 3, 6, 1, cases.genz_params("corner_peak", 6), [0] * 6, [1] * 6, 1e-12, NAN,
                      max_eval=3000000))
    out.append(record("oscvec_d1_f64_L2", 30, 1, 64, [], [0], [1], 1e-8, NAN, 2))
    out.append(record("c2_oscvec_d1_f1024_L2", 30, 1, 1024, [], [0], [1], 1e-8, NAN, 2))   # BASELINE config C2
    out.append(record("oscvec_d1_f7_LINF", 30, 1, 7, [], [0], [3], 1e-9, NAN, 4))
    out.append(record("oscvec_d2_f5_L1", 30, 2, 5, [], [0, 0], [1, 2], 1e-6, NAN, 3))
    out.append(record("cexp_d3_f8_paired_semiinf", 31, 3, 8, [0.5, 1.0, 2.0, 3.0], [0] * 3, [inf] * 3, 1e-4, NAN, 1,
                      max_eval=2000000, transform=[1, 1, 1]))
    out.append(record("gauss_d2_mixed_semiinf", 10, 2, 1, [0.7], [1.0, -inf], [inf, 0.5], 1e-6, NAN, 0,
                      max_eval=2000000, transform=[1, 2]))
    out.append(record("c4_cexp_d4_f32_paired", 31, 4, 32, [0.25 * (m + 1) for m in range(16)], [0] * 4, [inf] * 4, 1e-3, NAN, 1,
                      max_eval=600000, transform=[1, 1, 1, 1]))
    out.append(record("gauss_d2_inf", 10, 2, 1, [1.0], [-inf, -inf], [inf, inf], 1e-6, NAN, 0, max_eval=2000000, transform=[3, 3]))
    out.append(record("gauss_d1_inf", 10, 1, 1, [1.0], [-inf], [inf], 1e-9, NAN, 0, transform=[3]))
    with open(os.path.join(HERE, "oracle_pins.json"), "w") as f:
        json.dump(out, f, indent=1, allow_nan=False, default=lambda o: None)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
