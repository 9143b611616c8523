#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== bench default"; ( time timeout 1200 python bench.py > gpurun_out/r43_bench.txt 2>gpurun_out/r43_bench.err ) 2>&1 | grep real; tail -c 4500 gpurun_out/r43_bench.txt
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r43_ref.txt 2>&1; tail -c 700 gpurun_out/r43_ref.txt
echo "== sweep256"; timeout 900 python bench.py --workload sweep256 --no-cpu --no-cudnn --no-swap > gpurun_out/r43_sweep.txt 2>&1; grep -o '"sweep256": \[[^]]*\]' gpurun_out/r43_sweep.txt | head -c 2500
echo "== per-op"; timeout 600 python tools/profile_ops.py --repeat 20 --top 400 > gpurun_out/r43_per_op.txt 2>&1; grep "====" gpurun_out/r43_per_op.txt
echo "== elementwise bw"; timeout 300 python tools/elementwise_bw.py > gpurun_out/r43_elementwise_bw.txt 2>&1; tail -25 gpurun_out/r43_elementwise_bw.txt
