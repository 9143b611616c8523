#!/bin/bash
# head convolution on mma.sync / ldmatrix: tests, then A/B against the FFMA form (debug build knob D3FK_HEAD_MMA)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py -q -m gpu -p no:cacheprovider -k "head or eval or sampler or train_step or forward" 2>&1 | tail -3
A="-- --no-swap --no-cudnn --sample-steps 200"
bash tools/ab.sh "ffma D3FK_LIB=tools/libd3fk_dbg.so D3FK_HEAD_MMA=0 $A" "mma D3FK_LIB=tools/libd3fk_dbg.so D3FK_HEAD_MMA=1 $A" "product X=1 $A" 2>&1 | tee gpurun_out/r64_ab.txt
