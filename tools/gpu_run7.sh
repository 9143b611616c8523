#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== geometry"; D3FK_VERBOSE=1 D3FK_TRAIN_GRAPH=0 timeout 300 python tools/one_step_verbose.py > gpurun_out/r7_geometry.txt 2>&1; grep -c d3fk gpurun_out/r7_geometry.txt
echo "== split conv / bn"; timeout 600 python tools/split_convbn.py > gpurun_out/r7_split.txt 2>&1; cat gpurun_out/r7_split.txt
echo "== ncu"; bash tools/ncu_slab.sh "r7_s32 conv 262144 32 288 0" "r7_s32c128 conv 262144 32 1152 0" "r7_l1 conv 65536 64 576 0" "r7_l3 conv 4096 256 2304 0" "r7_t16 conv 1048576 16 288 0"
ls -la gpurun_out | grep r7_
