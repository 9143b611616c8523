#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-r24}
echo "== pytest all"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/${T}_pytest.txt 2>&1; tail -4 gpurun_out/${T}_pytest.txt
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-cudnn --no-swap --sample-steps 200 > gpurun_out/${T}_bench.txt 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/${T}_bench.txt
echo "== per-op"; timeout 600 python tools/profile_ops.py --repeat 20 --top 400 > gpurun_out/${T}_per_op.txt 2>&1; grep "====" gpurun_out/${T}_per_op.txt
