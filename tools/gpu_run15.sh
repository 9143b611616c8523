#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for spec in "4096 256 2304 0" "1024 512 4608 0"; do
  echo "== timeline $spec"; D3FK_LIB=tools/libd3fk_tl.so timeout 300 python tools/timeline.py $spec 2>&1 | tail -3
done
for cap in 100 200 50; do
echo "== split cap $cap"; D3FK_LIB=tools/libd3fk_dbg.so D3FK_SPLIT_CAP=$cap timeout 600 python tools/split_convbn.py 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7,$9,$10,$11}' | sort | uniq -c | sort -k2 | grep -v "N= 64 K= 576\|N= 16\|N= 32"
done
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-cudnn --no-swap --sample-steps 200 > gpurun_out/r15_bench.txt 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r15_bench.txt
