#!/bin/bash
# one-launch BN backward (grid barrier) re-measured with the lighter barrier (debug knob)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 D3FK_LIB=tools/libd3fk_dbg.so
A="-- --no-sample --no-swap --no-cudnn"
bash tools/ab.sh "dbg_base X=1 $A" "fuse_bnbwd D3FK_FUSE_BN_BWD=1 $A" "fuse_bnbwd_1M D3FK_FUSE_BN_BWD=1 D3FK_FUSE_BN_BWD_MAX=1100000 $A" \
  "fuse_bnbwd_17M D3FK_FUSE_BN_BWD=1 D3FK_FUSE_BN_BWD_MAX=17000000 $A" "dbg_base2 X=1 $A" 2>&1 | tee gpurun_out/r61_ab.txt
