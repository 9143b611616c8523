"""Per-kernel table from an `ncu --section SpeedOfLight --section SchedulerStats ... --csv --page raw` capture of the
tensor-core kernels of one training step: launches, total time, tensor-pipe active %, issue-slot utilisation ("No
Eligible" = 100 - issue active), L2 / DRAM throughput %, registers, shared memory.
usage: python tools/summarise_sections.py gpurun_out/conv_sections_r02.csv > profiles/conv_sections_r02_summary.txt"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]; col = {n: i for i, n in enumerate(hdr)}
body = [r for r in rows[hi + 2:] if len(r) == len(hdr)]
def short(n):
    n = re.sub(r'^void ', '', n); n = re.sub(r'\(.*', '', n)
    return n.replace('d3fk::', '')
f = lambda r, k: float(r[col[k]].replace(',', '') or 0)
agg = collections.OrderedDict()
for r in body:
    a = agg.setdefault(short(r[col["Kernel Name"]]), collections.defaultdict(float))
    t = f(r, "gpu__time_duration.sum") / 1e3
    a["n"] += 1; a["t"] += t
    for k, m in (("tensor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                 ("issue", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                 ("lts", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                 ("dram", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                 ("warps", "sm__warps_active.avg.per_cycle_active")):
        a[k] += t * f(r, m)            # time-weighted
    a["regs"] = f(r, "launch__registers_per_thread"); a["smem"] = max(a["smem"], f(r, "launch__shared_mem_per_block") / 1024)
tot = sum(a["t"] for a in agg.values())
print(f"# {len(body)} launches of the tensor-core kernels of one training step (B=256 @64x64, bf16), {tot:.0f} us cold-cache / serialised under ncu")
print(f"{'kernel':34s} {'launches':>8s} {'us':>8s} {'share':>6s} {'tensor pipe %':>13s} {'issue active %':>14s} {'no eligible %':>13s} {'L2 %':>6s} {'DRAM %':>6s} {'warps/SM':>8s} {'regs':>5s} {'smem KB':>8s}")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
    t = a["t"]
    print(f"{name:34s} {int(a['n']):8d} {t:8.1f} {100*t/tot:5.1f}% {a['tensor']/t:13.1f} {a['issue']/t:14.1f} {100-a['issue']/t:13.1f} {a['lts']/t:6.1f} {a['dram']/t:6.1f} {a['warps']/t:8.1f} {int(a['regs']):5d} {a['smem']:8.1f}")
