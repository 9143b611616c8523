#!/bin/bash
export PYTHONUNBUFFERED=1
D3FK_UPCAT_ALL=1 timeout 900 python -m pytest tests/test_gpu_unet.py -q -m gpu -p no:cacheprovider -x 2>&1 | grep -E "^E |Error|FAILED|passed|failed" | head -20
