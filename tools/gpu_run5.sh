#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest all"; timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r5_pytest.txt 2>&1; tail -8 gpurun_out/r5_pytest.txt
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-cudnn --sample-steps 200 > gpurun_out/r5_bench.txt 2>&1; tail -c 2600 gpurun_out/r5_bench.txt
echo "== per-op wgrad"; timeout 600 python tools/profile_ops.py --repeat 20 --top 400 > gpurun_out/r5_per_op.txt 2>&1; grep "OP_WGRAD " gpurun_out/r5_per_op.txt | sort -k1 -n -r | head -6
