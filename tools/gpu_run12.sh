#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
export D3FK_LIB=tools/libd3fk_tl.so
for spec in "4096 256 2304 0" "1024 512 4608 0" "4096 256 2304 1"; do
  echo "== timeline $spec"; timeout 300 python tools/timeline.py $spec 2>&1 | tail -6
done
