#!/bin/bash
# staged + coalesced epilogue of the slab weight gradient: tests, per-op times (debug build knobs), bench
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider -k "wgrad" 2>&1 | tail -4
for c in 0 1 2; do
  echo "== coalesce $c"; D3FK_LIB=tools/libd3fk_dbg.so D3FK_WGS_COALESCE=$c timeout 120 python tools/wgrad_slab_variants.py 2>&1 | grep -E "wgrad M|total"
done 2>&1 | tee gpurun_out/r58_wgs_coalesce.txt
A="-- --no-sample --no-swap --no-cudnn"
bash tools/ab.sh "new X=1 $A" "new2 X=1 $A" 2>&1 | tee gpurun_out/r58_ab.txt
