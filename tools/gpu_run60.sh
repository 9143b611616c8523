#!/bin/bash
# (1) lighter grid barrier (red.release arrive, no trailing fence / sleep): tests + A/B against the previous build
# (2) epilogue-store term in the weight-gradient split model (debug knob D3FK_WG_STORE_NS): per-op and step
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3
A="-- --no-swap --no-cudnn --sample-steps 200"
bash tools/ab.sh "old D3FK_LIB=tools/libd3fk_old.so $A" "new X=1 $A" "old2 D3FK_LIB=tools/libd3fk_old.so $A" "new2 X=1 $A" \
  "ns15 D3FK_LIB=tools/libd3fk_dbg.so D3FK_WG_STORE_NS=0.15 $A" "ns33 D3FK_LIB=tools/libd3fk_dbg.so D3FK_WG_STORE_NS=0.33 $A" 2>&1 | tee gpurun_out/r60_ab.txt
for ns in 0 0.33; do D3FK_LIB=tools/libd3fk_dbg.so D3FK_WG_STORE_NS=$ns timeout 120 python tools/wgrad_variants.py 2>&1 | grep -E "total|M= +1024|M= +4096" ; done | tee gpurun_out/r60_wg_ns.txt
