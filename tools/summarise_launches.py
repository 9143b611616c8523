"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time and share.
usage: python tools/summarise_launches.py gpurun_out/launches.csv > profiles/launches_summary.csv"""
import csv, io, re, sys, collections
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO(''.join(lines))))
agg = collections.OrderedDict()
tot = 0.0
for r in rows:
    n = r['Kernel Name']
    n = re.sub(r'^void ', '', n)
    n = re.sub(r'\(.*', '', n)
    n = n.replace('__nv_bfloat16', 'bf16')
    t = float(r['Metric Value']) / 1000.0
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += t
    tot += t
pk = [i for i, r in enumerate(rows) if 'pack_all' in r['Kernel Name']]
steps = len(pk) - 1 if len(pk) > 1 else 1
span = rows[pk[0]:pk[-1]] if len(pk) > 1 else rows
per_step = sum(float(r['Metric Value']) for r in span) / 1000.0 / steps
print(f"# {len(rows)} launches, {tot/1000:.3f} ms of kernel time; {steps} full training steps between pack_all launches: "
      f"{per_step/1000:.3f} ms and {len(span)//steps} launches per step (cold-cache, serialised: compare SHARES)")
print("kernel,launches,total_us,share")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n},{c},{t:.1f},{100*t/tot:.1f}%")
