#!/bin/bash
# ablation of the slab weight gradient (debug build; D3FK_WGS_ABLATE: 1 no atomics, 2 no MMAs, 4 no slab loads, 8 no dY load)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 D3FK_LIB=tools/libd3fk_dbg.so
for a in 0 1 2 3 4 8 12 14 15; do
  echo "== ablate $a"; D3FK_WGS_ABLATE=$a timeout 120 python tools/wgrad_slab_variants.py 2>&1 | grep -E "wgrad M|total"
done 2>&1 | tee gpurun_out/r57_wgs_ablate.txt
