#!/bin/bash
# which weight gradients are exposed?  debug build, D3FK_SKIP_WGRAD class mask (DROPS work: timing experiment only)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 D3FK_LIB=tools/libd3fk_dbg.so
A="-- --no-sample --no-swap --no-cudnn"
bash tools/ab.sh "dbg_base X=1 $A" "skip_all D3FK_SKIP_WGRAD=1 $A" "skip_M262k+ D3FK_SKIP_WGRAD=2 $A" "skip_M65k D3FK_SKIP_WGRAD=4 $A" \
  "skip_M16k D3FK_SKIP_WGRAD=8 $A" "skip_M4k D3FK_SKIP_WGRAD=16 $A" "skip_M1k D3FK_SKIP_WGRAD=32 $A" \
  "skip_enc D3FK_SKIP_WGRAD=60 $A" 2>&1 | tee gpurun_out/r55_ab.txt
