#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest all"; timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/r4_pytest.txt 2>&1; tail -15 gpurun_out/r4_pytest.txt
echo "== probe legacy (prefetch epilogue)"; D3FK_LIB=tools/libd3fk_dbg.so timeout 300 python tools/probe_slab_flat.py 2>&1 | grep " us "
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-cudnn --sample-steps 200 > gpurun_out/r4_bench.txt 2>&1; tail -c 2500 gpurun_out/r4_bench.txt
echo "== per-op"; timeout 600 python tools/profile_ops.py --repeat 20 --top 400 > gpurun_out/r4_per_op.txt 2>&1; grep -A16 "eval forward" gpurun_out/r4_per_op.txt
