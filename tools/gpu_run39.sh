#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python bench.py --steps 30 --warmup 5 --no-cpu --no-cudnn --no-swap --sample-steps 200"
run() { echo -n "$1: "; shift; env "$@" timeout 300 $B 2>&1 | grep -o '"ms_per_step": [0-9.]*' | tr '\n' ' '; echo; }
run base X=1
run upcat_all D3FK_UPCAT_ALL=1
run base X=1
run upcat_all D3FK_UPCAT_ALL=1
echo "== tests with upcat_all"; D3FK_UPCAT_ALL=1 timeout 900 python -m pytest tests/test_gpu_unet.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -2
echo "== per-op"; D3FK_UPCAT_ALL=1 timeout 600 python tools/profile_ops.py --repeat 20 --top 400 > gpurun_out/r39_per_op.txt 2>&1; grep "====" gpurun_out/r39_per_op.txt; grep "K=6912\|K=3456\|K=1728\|UPCAT" gpurun_out/r39_per_op.txt | head -30
