"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`):
per-kernel launches, total time, share and (when captured) DRAM bytes; steps are delimited by the q_sample launch that opens
every training step.   usage: python tools/summarise_launches.py gpurun_out/launches.csv > profiles/launches_summary.csv
With --json the conv-family totals per step are printed as JSON (profiles/conv_traffic_*.json, read by bench.py)."""
import csv, io, json, re, sys, collections
path = [a for a in sys.argv[1:] if not a.startswith("--")][0]
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO(''.join(lines))))
launches = collections.OrderedDict()          # ID -> {name, t_us, rd, wr}
for r in rows:
    L = launches.setdefault(r['ID'], {"name": r['Kernel Name'], "t": 0.0, "rd": 0.0, "wr": 0.0})
    v = float(r['Metric Value'].replace(',', ''))
    unit = r.get('Metric Unit', '')
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "nsecond": 1e-3}.get(unit, 1.0)
    m = r.get('Metric Name', 'gpu__time_duration.sum')
    if m.startswith('gpu__time_duration'):
        L["t"] += v * (scale if unit else 1e-3)
    elif m.startswith('dram__bytes_read'):
        L["rd"] += v * scale
    elif m.startswith('dram__bytes_write'):
        L["wr"] += v * scale
ls = list(launches.values())
def short(n):
    n = re.sub(r'^void ', '', n)
    n = re.sub(r'\(.*', '', n)
    return n.replace('__nv_bfloat16', 'bf16').replace('d3fk::', '')
starts = [i for i, L in enumerate(ls) if 'qsample_kernel' in L["name"]]
steps = max(1, len(starts) - 1)
span = ls[starts[0]:starts[-1]] if len(starts) > 1 else ls
is_conv = lambda n: any(k in n for k in ("conv_tc_kernel", "conv_slab_kernel", "wgrad_tc_kernel", "wgrad_slab_kernel", "head_conv_kernel"))
conv = [L for L in span if is_conv(L["name"])]
info = {"source": path, "steps": steps, "launches_per_step": len(span) / steps,
        "kernel_us_per_step": sum(L["t"] for L in span) / steps,
        "conv_family": {"launches_per_step": len(conv) / steps, "us_per_step": sum(L["t"] for L in conv) / steps,
                        "dram_read_bytes_per_step": sum(L["rd"] for L in conv) / steps,
                        "dram_write_bytes_per_step": sum(L["wr"] for L in conv) / steps},
        "all_kernels": {"dram_read_bytes_per_step": sum(L["rd"] for L in span) / steps,
                        "dram_write_bytes_per_step": sum(L["wr"] for L in span) / steps}}
if "--json" in sys.argv:
    print(json.dumps(info, indent=1))
    sys.exit(0)
agg = collections.OrderedDict()
for L in span:
    a = agg.setdefault(short(L["name"]), [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += L["t"]; a[2] += L["rd"]; a[3] += L["wr"]
tot = sum(a[1] for a in agg.values())
print(f"# {len(ls)} launches captured; {steps} full training steps (q_sample to q_sample): {info['kernel_us_per_step'] / 1e3:.3f} ms of "
      f"kernel time and {info['launches_per_step']:.0f} launches per step (cold-cache, serialised under ncu: compare SHARES)")
print(f"# conv family per step: {info['conv_family']['launches_per_step']:.0f} launches, {info['conv_family']['us_per_step'] / 1e3:.3f} ms, "
      f"DRAM read {info['conv_family']['dram_read_bytes_per_step'] / 1e6:.1f} MB + write {info['conv_family']['dram_write_bytes_per_step'] / 1e6:.1f} MB")
print("kernel,launches_per_step,us_per_step,share,dram_read_MB_per_step,dram_write_MB_per_step")
for n, (c, t, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n},{c / steps:.1f},{t / steps:.1f},{100 * t / tot:.1f}%,{rd / steps / 1e6:.1f},{wr / steps / 1e6:.1f}")
