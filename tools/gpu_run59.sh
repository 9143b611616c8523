#!/bin/bash
# weight-gradient co-residency knobs on the build with the coalesced slab epilogue (debug build)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 D3FK_LIB=tools/libd3fk_dbg.so
A="-- --no-sample --no-swap --no-cudnn"
bash tools/ab.sh "dbg_base X=1 $A" "occ1 D3FK_WG_OCC=1 $A" "prio D3FK_MAIN_PRIORITY=5 $A" "occ1_prio D3FK_WG_OCC=1 D3FK_MAIN_PRIORITY=5 $A" \
  "side2 D3FK_SIDE_STREAMS=2 $A" "side2_occ1 D3FK_SIDE_STREAMS=2 D3FK_WG_OCC=1 $A" "skip_tail D3FK_SKIP_WGRAD=2 $A" 2>&1 | tee gpurun_out/r59_ab.txt
