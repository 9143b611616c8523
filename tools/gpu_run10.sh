#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for fl in 0 1; do
  echo "== probe flat=$fl"; D3FK_LIB=tools/libd3fk_dbg.so D3FK_SLAB_FLAT=$fl timeout 300 python tools/probe_slab_flat.py > gpurun_out/r10_probe_flat$fl.txt 2>&1; grep -A14 "timing" gpurun_out/r10_probe_flat$fl.txt; grep -c "rel err" gpurun_out/r10_probe_flat$fl.txt; grep "FAILED\|error flag" gpurun_out/r10_probe_flat$fl.txt
done
for ab in 1 2 4 6; do
  echo "== flat=1 ablate=$ab"; D3FK_LIB=tools/libd3fk_dbg.so D3FK_SLAB_FLAT=1 D3FK_SLAB_ABLATE=$ab timeout 300 python tools/probe_slab_flat.py 2>&1 | grep " us "
done
for ab in 1 2 4 6; do
  echo "== flat=0 ablate=$ab"; D3FK_LIB=tools/libd3fk_dbg.so D3FK_SLAB_FLAT=0 D3FK_SLAB_ABLATE=$ab timeout 300 python tools/probe_slab_flat.py 2>&1 | grep " us "
done
