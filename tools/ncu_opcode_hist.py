"""Dynamic opcode histogram of one kernel from an `ncu --page source --csv` export (full, or the trimmed copies under
profiles/): warp-level instructions per executed warp-tile, FP64 share, and the issue-slot bound it implies
(2 issue cycles per warp-wide FP64 instruction, 1 per other instruction: tools/issue_probe.cu).

usage: python tools/ncu_opcode_hist.py <source.csv> [once_per_tile_opcode=LDGDEPBAR|MUFU...]"""
import csv
import sys
from collections import Counter


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    hdr, data = rows[idx[-1]], rows[idx[-1] + 1:]
    col = {h: i for i, h in enumerate(hdr)}
    n, samples = Counter(), Counter()
    execs = []
    for r in data:
        if len(r) <= col["Instructions Executed"]:
            continue
        s = r[col["Source"]].split()
        if not s:
            continue
        op = (s[1] if s[0].startswith("@") else s[0]).rstrip(";").split(".")[0]
        e = int(r[col["Instructions Executed"]] or 0)
        n[op] += e
        samples[op] += int(r[col["# Samples"]] or 0)
        if e:
            execs.append(e)
    tot = sum(n.values())
    # warp-tiles: the most common execution count among the straight-line instructions of the tile body
    tiles = Counter(execs).most_common(1)[0][0]
    fp64 = sum(v for k, v in n.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    print("warp-tiles %d, warp instructions per tile %.1f, FP64 %.1f (%.1f %%), other %.1f" % (tiles, tot / tiles, fp64 / tiles, 100.0 * fp64 / tot, (tot - fp64) / tiles))
    print("issue-slot bound of the FP64 pipe: 2 N64 / (2 N64 + Nx) = %.3f" % (2.0 * fp64 / (2.0 * fp64 + (tot - fp64))))
    print("| opcode | per warp-tile | % of instructions | % of stall samples |\n|---|---|---|---|")
    ts = sum(samples.values()) or 1
    for op, v in n.most_common(24):
        print("| %s | %.1f | %.1f | %.1f |" % (op, v / tiles, 100.0 * v / tot, 100.0 * samples[op] / ts))


if __name__ == "__main__":
    main()
