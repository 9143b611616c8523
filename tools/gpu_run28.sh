#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python bench.py --steps 30 --warmup 5 --no-cpu --no-cudnn --no-swap --no-sample"
run() { echo -n "$1: "; shift; env "$@" timeout 300 $B 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1; }
run base X=1
run group2 D3FK_WGRAD_GROUP=1 D3FK_WGRAD_GROUP_SIZE=2
run base_dbg D3FK_LIB=tools/libd3fk_dbg.so
run base X=1
run group2 D3FK_WGRAD_GROUP=1 D3FK_WGRAD_GROUP_SIZE=2
run group2_dbg D3FK_LIB=tools/libd3fk_dbg.so D3FK_WGRAD_GROUP=1 D3FK_WGRAD_GROUP_SIZE=2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
