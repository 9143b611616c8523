#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-r32}
echo "== pytest all (-s for the parity metrics)"; timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/${T}_pytest.txt 2>&1; tail -4 gpurun_out/${T}_pytest.txt
grep "x0_hat rel err\|arena:\|bucket\|PSNR\|per-tensor\|passed\|failed" gpurun_out/${T}_pytest.txt > gpurun_out/${T}_parity.txt
echo "== bench"; timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu --no-cudnn --no-swap --sample-steps 200 > gpurun_out/${T}_bench.txt 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/${T}_bench.txt
