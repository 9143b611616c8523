#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest unet+ops"; timeout 1800 python -m pytest tests/test_gpu_unet.py tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider > gpurun_out/r36_pytest.txt 2>&1; tail -4 gpurun_out/r36_pytest.txt
