"""Aggregate an `ncu --page source --csv` export into contiguous regions of equal per-instruction execution count
(= loop bodies / roles): instructions, warp-level executions, stall samples and the dominant stall reasons per region."""
import csv, gzip, sys
path = sys.argv[1]
rows = list(csv.reader(gzip.open(path, "rt") if path.endswith(".gz") else open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]; col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
tot = sum(int(r[col["# Samples"]] or 0) for r in body)
print("kernel:", rows[0][1][:100], "samples", tot, "instructions", len(body))
def flush(start, end, ex, n, samp, agg):
    if n == 0: return
    top = ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:4] if v)
    print(f"[{start:5d}-{end:5d}] n={n:5d} exec/instr~{ex:8d} warp-instr={ex*n:10d} samples={samp:5d} ({100*samp/max(tot,1):4.1f}%) {top}")
start = 0; cur = None; n = 0; samp = 0; agg = {}
for i, r in enumerate(body):
    ex = int(r[col["Instructions Executed"]] or 0)
    # new region when the execution count changes by more than 2x
    if cur is None or ex > 2 * cur or cur > 2 * max(ex, 1):
        if cur is not None: flush(start, i - 1, cur, n, samp, agg)
        start = i; cur = ex; n = 0; samp = 0; agg = {}
    n += 1; samp += int(r[col["# Samples"]] or 0)
    for s in stalls: agg[s] = agg.get(s, 0) + int(r[col[s]] or 0)
flush(start, len(body) - 1, cur, n, samp, agg)
