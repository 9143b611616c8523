#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-r38}
echo "== pytest ops+unet+elementwise"; timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py tests/test_gpu_elementwise.py -q -m gpu -p no:cacheprovider -x > gpurun_out/${T}_pytest.txt 2>&1; tail -3 gpurun_out/${T}_pytest.txt
echo "== bench"; timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu --no-cudnn --no-swap --sample-steps 200 > gpurun_out/${T}_bench.txt 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/${T}_bench.txt
echo "== per-op"; timeout 600 python tools/profile_ops.py --repeat 20 --top 400 > gpurun_out/${T}_per_op.txt 2>&1; grep "====" gpurun_out/${T}_per_op.txt; grep -A6 "==== pack" gpurun_out/${T}_per_op.txt | tail -5
