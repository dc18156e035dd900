"""Top stall locations from `ncu --page source --csv` output: python tools/ncu_hot.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if "Source" in r)
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for n, r in enumerate(rows[hi + 1:]):
    if len(r) < len(hdr):
        continue
    data.append((int(r[idx["# Samples"]] or 0), n, r))
total = sum(d[0] for d in data)
print("total samples", total)
agg = {s: sum(int(d[2][idx[s]] or 0) for d in data) for s in stalls}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for smp, n, r in sorted(data, key=lambda d: -d[0])[:top]:
    why = {s[6:]: int(r[idx[s]] or 0) for s in stalls if int(r[idx[s]] or 0)}
    print(f"{smp:6d} {smp/total*100:5.1f}%  #{n:4d} {r[idx['Source']].strip()[:70]:70s} {why}")
