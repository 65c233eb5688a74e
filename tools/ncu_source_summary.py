#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` of ONE kernel: stall reasons, instruction mix, samples per opcode and the
hottest instructions. usage: ncu -i rep --page source --csv --kernel-name regex:NAME --launch-count 1 | python tools/ncu_source_summary.py"""
import csv, collections, re, sys
rows = [r for r in csv.reader(sys.stdin)]
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; idx = {n: i for i, n in enumerate(h)}
stall = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = collections.Counter(); byop = collections.Counter(); execs = collections.Counter(); samples = 0; hot = []
seen = set()
for r in rows[hi + 1:]:
    if len(r) < len(h) or r[0] == "Address": continue
    if r[0] in seen: continue           # the page lists every instruction twice (SASS view and source-correlated view)
    seen.add(r[0])
    src = r[idx["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else "?"
    base = op.split(".")[0]
    if base in ("MUFU", "LDS", "LDG", "STG", "STS", "SHFL"): base = op
    ex = int(r[idx["Instructions Executed"]] or 0); s = int(r[idx["# Samples"]] or 0)
    execs[base] += ex; samples += s; byop[base] += s
    for n in stall: tot[n] += int(r[idx[n]] or 0)
    hot.append((s, src[:70], {n[6:]: int(r[idx[n]] or 0) for n in stall if int(r[idx[n]] or 0) > s * 0.25 and s > 0}))
print(rows[0][1][:100] if rows and len(rows[0]) > 1 else "")
print("samples", samples, "warp instructions", sum(execs.values()))
print("stall reasons:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / samples) for k, v in tot.most_common(10)))
te = sum(execs.values())
print("instruction mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / te) for k, v in execs.most_common(22)))
print("samples by opcode:", ", ".join("%s %.1f%%" % (k, 100.0 * v / samples) for k, v in byop.most_common(14)))
print("hottest instructions:")
for s, src, why in sorted(hot, key=lambda t: -t[0])[:14]:
    print("  %5.2f%%  %-70s %s" % (100.0 * s / samples, src, why))
