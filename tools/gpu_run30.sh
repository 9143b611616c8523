#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
export D3FK_LIB=tools/libd3fk_dbg.so
for tma in 1 0 1 0; do
  echo "== TMA=$tma"; D3FK_WG_TMA=$tma timeout 600 python -m pytest tests/test_gpu_parity_configs.py -q -m gpu -p no:cacheprovider -s -k "test_train_step_gradients_benchmarked_configs and bf16" 2>&1 | grep "bucket 1\|arena\|passed\|failed"
done
