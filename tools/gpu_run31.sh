#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for i in 1 2 3; do
  echo "== run $i"; timeout 900 python -m pytest tests/test_gpu_parity_configs.py -q -m gpu -p no:cacheprovider -s -k "not two_gpus" > gpurun_out/r31_parity_$i.txt 2>&1; grep "arena\|passed\|failed\|x0_hat\|PSNR\|psnr" gpurun_out/r31_parity_$i.txt | cut -c1-160
done
