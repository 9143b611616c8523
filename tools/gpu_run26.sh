#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python bench.py --steps 30 --warmup 5 --no-cpu --no-cudnn --no-swap --no-sample"
export D3FK_LIB=tools/libd3fk_dbg.so
run() { echo -n "$1: "; shift; env "$@" timeout 300 $B 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1; }
run base X=1
run base2 X=1
run wg_occ1 D3FK_WG_OCC=1
run wg_slab0 D3FK_WG_SLAB=0
run lane0 D3FK_BRANCH_LANE=0
run fuse_bnbw0 D3FK_FUSE_BNBW=0
run fuse_bnbw_all D3FK_FUSE_BNBW=1
run fuse_bnbw_4M D3FK_FUSE_BNBW=4500000
run split_cap200 D3FK_SPLIT_CAP=200
run wg_occ1_slab0 D3FK_WG_OCC=1 D3FK_WG_SLAB=0
