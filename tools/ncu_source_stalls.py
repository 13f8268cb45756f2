#!/usr/bin/env python
"""Where a kernel's warp-stall samples sit: python tools/ncu_source_stalls.py prof.ncu-rep [N]
prints the N SASS instructions with the most stall samples (with the dominant reasons) and a cumulative profile along the program."""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ci["# Samples"]] or 0) for r in data)
print(f"# {rows[0][1][:90]}: {len(data)} instructions, {tot} samples")
agg = {}
for r in data:
    for h in stall_cols:
        agg[h] = agg.get(h, 0) + int(r[ci[h]] or 0)
print("by reason: " + ", ".join(f"{h[6:]} {v/tot:.1%}" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(data)), key=lambda i: -int(data[i][ci["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = data[i]; n = int(r[ci["# Samples"]] or 0)
    why = sorted(((int(r[ci[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {n/tot:6.1%}  {r[ci['Source']].strip()[:70]:70s} {why[0][1]} {why[0][0]} {why[1][1]} {why[1][0]}")
# cumulative by tenths of the program
step = max(1, len(data) // 20)
print("cumulative share of samples along the program (5 % steps): " +
      " ".join(f"{sum(int(r[ci['# Samples']] or 0) for r in data[k:k+step])/tot:.0%}" for k in range(0, len(data), step)))
