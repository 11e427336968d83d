#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from an ncu report's source page.
usage: ncu_hot.py <file.ncu-rep> [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = rows[hdr_i + 1:]
total = sum(int(r[col["# Samples"]] or 0) for r in data)
print(f"total samples {total}")
ranked = sorted(enumerate(data), key=lambda ir: -int(ir[1][col["# Samples"]] or 0))[:top]
for idx, r in ranked:
    s = int(r[col["# Samples"]] or 0)
    reasons = sorted(((int(r[col[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{idx:6d} {s:6d} {100.0 * s / max(1, total):5.1f}%  {r[col['Source']].strip()[:70]:70s} " +
          " ".join(f"{n}:{c}" for c, n in [(c, n) for n, c in reasons] if n))
