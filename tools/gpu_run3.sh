#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python bench.py --steps 20 --warmup 5 --no-cpu --no-cudnn --no-swap --no-sample"
export D3FK_LIB=tools/libd3fk_dbg.so
echo "== base";            timeout 300 $B 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
echo "== skip wgrad";      D3FK_SKIP_WGRAD=1 timeout 300 $B 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
echo "== no fork (serial)"; D3FK_FORK_WGRAD=0 timeout 300 $B 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
echo "== 1 side stream";   D3FK_SIDE_STREAMS=1 timeout 300 $B 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
echo "== timeline skip wgrad"; D3FK_TRAIN_GRAPH=0 D3FK_SKIP_WGRAD=1 timeout 300 python tools/step_timeline.py 2>&1 | tail -12
echo "== timeline base"; D3FK_TRAIN_GRAPH=0 timeout 300 python tools/step_timeline.py 2>&1 | tail -12
for ab in 0 1 2 4 3 5 6; do
  echo "== slab ablate=$ab"; D3FK_SLAB_FLAT=0 D3FK_SLAB_ABLATE=$ab timeout 300 python tools/probe_slab_flat.py 2>&1 | grep " us " 
done
echo "== grad profile"; timeout 600 python tools/diag_grad_profile.py 256 21 23 > gpurun_out/r3_grad_profile.txt 2>&1; head -80 gpurun_out/r3_grad_profile.txt
unset D3FK_LIB
echo "== pytest parity (shipped lib)"; timeout 900 python -m pytest tests/test_gpu_parity_configs.py -q -m gpu -s -p no:cacheprovider -k "gradients" > gpurun_out/r3_pytest_grad.txt 2>&1; grep -v "^$" gpurun_out/r3_pytest_grad.txt | grep -v "^tests/\|^  \|Warning\|warnings" | tail -60
