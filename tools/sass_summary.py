"""SASS summary of the NVRTC-built scalar rule kernel (cub_gm_scalar) of one integrand: opcode histogram, FP64 instruction
mix, local-memory sites.  Needs no GPU (NVRTC cross-compiles for sm_100a).
    python tools/sass_summary.py corner_peak 6 > profiles/r02/sass_gm_scalar_corner_peak_d6.md"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cubature_b200 as cb  # noqa: E402
import cases  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "corner_peak"
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ig = cb.Integrand.builtin(cases.GENZ[name], dim, 1, cases.genz_params(name, dim))
with tempfile.TemporaryDirectory() as td:
    path = os.path.join(td, "k.cubin")
    open(path, "wb").write(ig.cubin())
    res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "cub_gm_scalar", path], capture_output=True, text=True).stdout
lines = [l for l in sass.splitlines() if re.search(r"/\*[0-9a-f]{4,5}\*/\s", l)]
ops = collections.Counter()
local = []
for l in lines:
    m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)\s*(.*?);", l)
    if not m:
        continue
    op = m.group(3)
    ops[op] += 1
    if op in ("LDL", "STL"):
        local.append((m.group(1), op + m.group(4), m.group(5).strip()))
tot = sum(ops.values())
fp64 = {k: ops[k] for k in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU")}
nfp = ops["DFMA"] + ops["DMUL"] + ops["DADD"]
print("# SASS of `cub_gm_scalar`, Genz %s, dim %d (NVRTC, sm_100a, -fmad=false)\n" % (name, dim))
for l in res.splitlines():
    if "cub_gm_scalar" in l or ("REG:" in l and "cub_gm_scalar" in prev):
        print("    " + l.strip())
    prev = l
print("\n%d instructions (static).  FP64 arithmetic: DFMA %d, DMUL %d, DADD %d = %d (%.1f %% of all); FMA share of the FP64 arithmetic %.1f %%;"
      " MUFU (reciprocal seeds) %d, DSETP %d.\n" % (tot, ops["DFMA"], ops["DMUL"], ops["DADD"], nfp, 100.0 * nfp / tot, 100.0 * ops["DFMA"] / max(1, nfp),
                                                   ops["MUFU"], ops["DSETP"]))
print("| opcode | count | share |\n|---|---|---|")
for k, v in ops.most_common(24):
    print("| %s | %d | %.1f %% |" % (k, v, 100.0 * v / tot))
print("\nLocal-memory sites: %d STL, %d LDL (stack frame in the resource line above).\n" % (sum(1 for x in local if x[1].startswith("STL")),
                                                                                         sum(1 for x in local if x[1].startswith("LDL"))))
print("| address | instruction | operands |\n|---|---|---|")
for a, o, r in local:
    print("| 0x%s | %s | `%s` |" % (a, o, r))
