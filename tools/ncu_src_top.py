"""Top SASS lines by stall samples from `ncu --page source --csv` output (first kernel only)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
# find header row
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address" or r[0].startswith("Kernel"):
        if r and r[0].startswith("Kernel"): break
        continue
    try:
        data.append((int(r[idx["# Samples"]] or 0), int(r[idx["Instructions Executed"]] or 0), r))
    except ValueError:
        pass
tot = sum(d[0] for d in data); toti = sum(d[1] for d in data)
print("instructions", len(data), "samples", tot, "inst executed", toti)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(d[2][idx[h]] or 0) for d in data) for h in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for s, ie, r in sorted(data, key=lambda d: -d[0])[:n]:
    top = sorted(((int(r[idx[h]] or 0), h) for h in stalls), reverse=True)[:2]
    print(f"{s:7d} {100*s/tot:5.1f}%  exec {ie:10d}  {r[idx['Source']][:90]:90s} {top}")
