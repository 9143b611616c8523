#!/bin/bash
# weight-gradient kernels on a SUBSET of the SMs (smaller grids, longer kernels, SMs left free for the main chain)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 D3FK_LIB=tools/libd3fk_dbg.so
A="-- --no-sample --no-swap --no-cudnn"
bash tools/ab.sh "dbg_base X=1 $A" "cap148 D3FK_WG_CAP=148 $A" "cap96 D3FK_WG_CAP=96 $A" "cap64 D3FK_WG_CAP=64 $A" \
  "wgs111 D3FK_WGS_GRID=111 $A" "wgs74 D3FK_WGS_GRID=74 $A" "cap148_wgs111 D3FK_WG_CAP=148 D3FK_WGS_GRID=111 $A" \
  "cap96_wgs74 D3FK_WG_CAP=96 D3FK_WGS_GRID=74 $A" 2>&1 | tee gpurun_out/r63_ab.txt
