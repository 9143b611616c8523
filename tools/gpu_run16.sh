#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for ns in 0 1; do
for spec in "4096 256 2304 0" "16384 128 1152 0"; do
  echo "== timeline $spec nostats=$ns"; NOSTATS=$ns D3FK_LIB=tools/libd3fk_tl.so timeout 300 python tools/timeline.py $spec 2>&1 | tail -2
done
done
