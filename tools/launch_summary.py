"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total and share."""
import collections
import csv
import sys


def load(path):
    hdr, seq = None, []
    for r in csv.reader(open(path, errors="replace")):
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            u = d["Metric Unit"]
            v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else (v * 1e6 if u in ("s", "second") else v))
            seq.append((d["Kernel Name"].split("(")[0][:40], d["Grid Size"], v))
    return seq


if __name__ == "__main__":
    seq = load(sys.argv[1])
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(seq)
    seq = seq[lo:hi]
    agg = collections.OrderedDict()
    for k, g, v in seq:
        a = agg.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
    tot = sum(a[1] for a in agg.values())
    print("launches %d..%d of %s: %.1f us in kernels" % (lo, hi, sys.argv[1], tot))
    print("| kernel | launches | total us | share | mean us | max us |\n|---|---|---|---|---|---|")
    for k, a in agg.items():
        print("| %s | %d | %.1f | %.1f %% | %.2f | %.1f |" % (k, a[0], a[1], 100 * a[1] / tot, a[1] / a[0], a[2]))
