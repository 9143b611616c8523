#!/bin/bash
# grouped weight-gradient launch variants against the final build (the TMA-fed single launches are the default)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
A="-- --no-sample --no-swap --no-cudnn"
bash tools/ab.sh "base X=1 $A" "group D3FK_WGRAD_GROUP=1 $A" "group2 D3FK_WGRAD_GROUP=1 D3FK_WGRAD_GROUP_SIZE=2 $A" \
  "group3 D3FK_WGRAD_GROUP=1 D3FK_WGRAD_GROUP_SIZE=3 $A" "group4 D3FK_WGRAD_GROUP=1 D3FK_WGRAD_GROUP_SIZE=4 $A" \
  "group6 D3FK_WGRAD_GROUP=1 D3FK_WGRAD_GROUP_SIZE=6 $A" "base2 X=1 $A" 2>&1 | tee gpurun_out/r54_ab.txt
