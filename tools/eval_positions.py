#!/usr/bin/env python3
"""Evaluation positions of an adaptive integration, from a pool checkpoint (no GPU needed).

The reference's examples/PlotEvalPositions.java:28-54 records every point its integrand is called at and scatter-plots
them.  Here the integrand runs on the device, so the positions are reconstructed instead: a pool checkpoint
(cub_options_t.checkpoint_save) holds the final partition, and the rule's points in each box are a pure function of the box
(cubature_cuda_rule_points: Rule75GenzMalik.java:237-326 / RuleGaussKronrod_1d.java:60-84).  Coordinates are those of the
integration variable t (for transformed dimensions x = shift +- t/(1-t), README.md:237-272).

    python tools/eval_positions.py pool.ckpt positions.csv      # one row per point: region, x_0 ... x_{dim-1}
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import cubature_b200 as cb  # noqa: E402


def eval_positions(path, first=0, count=None):
    """-> (points [n * P, dim], region index of every point [n * P])"""
    centers, halfw, _, _, _ = cb.checkpoint_regions(path, first, count)
    n = len(centers)
    if n == 0:
        return np.zeros((0, cb.checkpoint_info(path)["dim"])), np.zeros(0, dtype=np.int64)
    pts = np.concatenate([cb.rule_points(centers[i], halfw[i]) for i in range(n)])
    return pts, np.repeat(np.arange(first, first + n), len(pts) // n)


if __name__ == "__main__":
    if len(sys.argv) < 3:
        raise SystemExit(__doc__)
    info = cb.checkpoint_info(sys.argv[1])
    pts, reg = eval_positions(sys.argv[1])
    np.savetxt(sys.argv[2], np.column_stack([reg, pts]), delimiter=",", fmt=["%d"] + ["%.17g"] * info["dim"],
               header="region," + ",".join("x%d" % d for d in range(info["dim"])))
    print("%d regions, %d points (%d per region) -> %s" % (info["n_regions"], len(pts), len(pts) // max(1, info["n_regions"]), sys.argv[2]))
