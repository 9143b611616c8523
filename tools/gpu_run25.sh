#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-r25}
echo "== pytest ops+unet"; timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py -q -m gpu -p no:cacheprovider -x > gpurun_out/${T}_pytest.txt 2>&1; tail -3 gpurun_out/${T}_pytest.txt
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-cudnn --no-swap --sample-steps 200 > gpurun_out/${T}_bench.txt 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/${T}_bench.txt
echo "== split"; timeout 600 python tools/split_convbn.py > gpurun_out/${T}_split.txt 2>&1; grep "total\|N= 128 K= 1152\|N= 256 K= 2304\|N= 512 K= 4608" gpurun_out/${T}_split.txt | sort | uniq -c | sort -rn | head -8
