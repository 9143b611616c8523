#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python bench.py --steps 20 --warmup 5 --no-cpu --no-cudnn --no-swap --no-sample"
export D3FK_LIB=tools/libd3fk_dbg.so
echo "== base";            timeout 300 $B 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
echo "== skip wgrad";      D3FK_SKIP_WGRAD=1 timeout 300 $B > gpurun_out/r22_skip.txt 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r22_skip.txt | head -1; tail -3 gpurun_out/r22_skip.txt | cut -c1-300
echo "== no fork (serial)"; D3FK_FORK_WGRAD=0 timeout 300 $B 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
echo "== eager (no graph)"; D3FK_TRAIN_GRAPH=0 timeout 300 $B 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
echo "== timeline base"; D3FK_TRAIN_GRAPH=0 timeout 300 python tools/step_timeline.py 2>&1 | tail -12
echo "== timeline skip wgrad"; D3FK_TRAIN_GRAPH=0 D3FK_SKIP_WGRAD=1 timeout 300 python tools/step_timeline.py 2>&1 | tail -12
