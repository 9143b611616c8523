#!/bin/bash
# A/B bench runs: each argument is "label ENV=VAL ... [-- bench args]"; prints ms/step, img/s, e2e, launches, conv ms, sample ms
for spec in "$@"; do
  label=${spec%% *}; rest=${spec#* }
  envs=${rest%%--*}; args=""
  case "$rest" in *--*) args=${rest#*--};; esac
  env $envs timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu $args > gpurun_out/ab_$label.log 2>&1
  python - "$label" <<'PY'
import json, sys
label = sys.argv[1]
ok = False
for l in open(f"gpurun_out/ab_{label}.log"):
    if l.startswith("{"):
        d = json.loads(l); ok = True
        s = d.get("sample") or {}
        print(f"{label:28s} {d['ms_per_step']:.3f} ms  {d['value']:.0f} img/s  e2e {d['e2e']['value']:.0f}  launches {d['gpu_launches']}  conv {d['roofline']['ms_per_step']:.3f}  sample {s.get('ms_per_step', 0):.3f}  loss {d['final_loss']:.4f}")
if not ok:
    print(label, "FAILED"); print(open(f"gpurun_out/ab_{label}.log").read()[-600:])
PY
done
