#!/usr/bin/env python
"""Per-region (source-line ranges) shares and SASS opcode mix of an .ncu-rep source page.

    python tools/ncu_regions.py <report.ncu-rep> <file> <name:lo-hi> [<name:lo-hi> ...]
"""
import collections, csv, io, subprocess, sys

rep, fname = sys.argv[1], sys.argv[2]
regions = []
for a in sys.argv[3:]:
    n, r = a.split(":")
    lo, hi = r.split("-")
    regions.append((n, int(lo), int(hi)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))


def num(x):
    try:
        return int(x)
    except Exception:
        return 0


cur, hdr, curline = None, None, None
agg = collections.defaultdict(lambda: [0, 0])
ops = collections.defaultdict(collections.Counter)
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[0] == "Function Name":
        continue
    d = dict(zip(hdr, r))
    if r[0].isdigit():
        curline = (cur, int(r[0]))
        agg[curline][0] += num(d.get("# Samples"))
        agg[curline][1] += num(d.get("Instructions Executed"))
    elif curline is not None and len(r) > 3 and r[3] not in ("-", ""):
        ins = r[3].split()
        if not ins:
            continue
        op = ins[1] if ins[0].startswith("@") and len(ins) > 1 else ins[0]
        ops[curline][op.split(".")[0]] += num(d.get("Instructions Executed"))
tots = sum(v[0] for v in agg.values()) or 1
toti = sum(v[1] for v in agg.values()) or 1


def region(f, ln):
    if f == fname:
        for n, lo, hi in regions:
            if lo <= ln <= hi:
                return n
        return fname + " (other)"
    return f


g = collections.defaultdict(lambda: [0, 0])
gops = collections.defaultdict(collections.Counter)
for k, v in agg.items():
    rg = region(*k)
    g[rg][0] += v[0]
    g[rg][1] += v[1]
    gops[rg].update(ops.get(k, {}))
print(f"total warp instructions {toti}, samples {tots}")
for k, v in sorted(g.items(), key=lambda x: -x[1][1]):
    mix = ", ".join(f"{o} {n / max(v[1], 1):.2f}" for o, n in gops[k].most_common(8))
    print(f"{k:26s} inst {100 * v[1] / toti:5.1f}%  samples {100 * v[0] / tots:5.1f}%   {mix}")
