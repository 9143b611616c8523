"""Top SASS instructions of an `ncu --page source --csv` export by warp-stall samples, with the dominant stall reason."""
import csv, gzip, sys
path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(gzip.open(path, "rt") if path.endswith(".gz") else open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
tot = sum(int(r[col["# Samples"]] or 0) for r in body)
print("kernel:", rows[0][1][:120], " total samples:", tot, " instructions:", len(body))
agg = {}
for r in body:
    for s in stalls:
        agg[s] = agg.get(s, 0) + int(r[col[s]] or 0)
print("stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
idx = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]] or 0))[:top]
for i in sorted(idx):
    r = body[i]
    n = int(r[col["# Samples"]] or 0)
    dom = max(stalls, key=lambda s: int(r[col[s]] or 0))
    print(f"{i:5d} {n:6d} {100.0 * n / max(tot, 1):5.1f}%  {dom[6:]:14s} exec={r[col['Instructions Executed']]:>8s}  {r[col['Source']].strip()[:110]}")
