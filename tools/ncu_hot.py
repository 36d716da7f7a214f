#!/usr/bin/env python
"""Top stall sites from an `ncu --page source --csv` export (SASS view): prints the N instructions with most samples
and the per-reason totals."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
iS, iSamp, iEx = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ridx = {r: hdr.index(r) for r in reasons}
data = []
tot = {r: 0 for r in reasons}
total = 0
for k, r in enumerate(rows[2:]):
    if len(r) < len(hdr): continue
    if not (r[iSamp] or '0').isdigit(): continue
    s = int(r[iSamp] or 0)
    total += s
    for q in reasons: tot[q] += int(r[ridx[q]] or 0)
    data.append((s, k, r))
print("total samples", total)
print("by reason:", ", ".join(f"{q[6:]}={v} ({100*v/max(total,1):.1f}%)" for q, v in sorted(tot.items(), key=lambda x: -x[1]) if v))
for s, k, r in sorted(data, key=lambda x: -x[0])[:N]:
    top = sorted(((int(r[ridx[q]] or 0), q[6:]) for q in reasons), reverse=True)[:3]
    print(f"{s:7d} {100*s/total:5.1f}%  #{k:5d} ex={r[iEx]:>8s}  {r[iS].strip()[:70]:70s} {top}")
