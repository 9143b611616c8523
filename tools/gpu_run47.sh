#!/bin/bash
export PYTHONUNBUFFERED=1
echo "== timeline base (eager)"; D3FK_TRAIN_GRAPH=0 timeout 300 python tools/step_timeline.py 2>&1 | tail -12
echo "== timeline skip wgrad (debug lib)"; D3FK_LIB=tools/libd3fk_dbg.so D3FK_TRAIN_GRAPH=0 D3FK_SKIP_WGRAD=1 timeout 300 python tools/step_timeline.py 2>&1 | tail -12
