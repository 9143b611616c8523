#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest all"; ( time timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r11_pytest.txt 2>&1 ) 2>&1 | grep real; tail -6 gpurun_out/r11_pytest.txt
echo "== bench default"; ( time timeout 1200 python bench.py > gpurun_out/r11_bench.txt 2>gpurun_out/r11_bench.err ) 2>&1 | grep real; tail -c 6000 gpurun_out/r11_bench.txt; tail -5 gpurun_out/r11_bench.err
echo "== bench reference arm"; ( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r11_ref.txt 2>&1 ) 2>&1 | grep real; tail -c 1500 gpurun_out/r11_ref.txt
echo "== sweep256"; ( time timeout 900 python bench.py --workload sweep256 --no-cpu > gpurun_out/r11_sweep.txt 2>&1 ) 2>&1 | grep real; tail -c 3000 gpurun_out/r11_sweep.txt
