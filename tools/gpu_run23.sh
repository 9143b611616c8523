#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python bench.py --steps 20 --warmup 5 --no-cpu --no-cudnn --no-swap --no-sample"
echo "== base";            timeout 300 $B 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
echo "== group wgrad";      D3FK_WGRAD_GROUP=1 timeout 300 $B 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
echo "== group wgrad, 1 side stream";      D3FK_SIDE_STREAMS=1 D3FK_WGRAD_GROUP=1 timeout 300 $B 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
echo "== 1 side stream";      D3FK_SIDE_STREAMS=1 timeout 300 $B 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
echo "== 6 side streams";      D3FK_SIDE_STREAMS=6 timeout 300 $B 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
echo "== per-op (all)"; timeout 600 python tools/profile_ops.py --repeat 20 --top 400 > gpurun_out/r23_per_op.txt 2>&1; grep "====" gpurun_out/r23_per_op.txt
echo "== per-op group"; D3FK_WGRAD_GROUP=1 timeout 600 python tools/profile_ops.py --repeat 20 --top 400 > gpurun_out/r23_per_op_group.txt 2>&1; grep "====\|WGRAD_GROUP" gpurun_out/r23_per_op_group.txt
