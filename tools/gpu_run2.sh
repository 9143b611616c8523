#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for bo in 1 0 2; do
  echo "== probe flat bo=$bo"; D3FK_LIB=tools/libd3fk_dbg.so D3FK_SLAB_FLAT=1 D3FK_SLAB_BO=$bo timeout 300 python tools/probe_slab_flat.py > gpurun_out/r2_probe_flat_bo$bo.txt 2>&1; tail -30 gpurun_out/r2_probe_flat_bo$bo.txt
done
echo "== probe legacy"; D3FK_LIB=tools/libd3fk_dbg.so D3FK_SLAB_FLAT=0 timeout 300 python tools/probe_slab_flat.py > gpurun_out/r2_probe_legacy.txt 2>&1; tail -16 gpurun_out/r2_probe_legacy.txt
echo "== pytest parity configs (legacy slab)"; D3FK_LIB=tools/libd3fk_dbg.so D3FK_SLAB_FLAT=0 timeout 900 python -m pytest tests/test_gpu_parity_configs.py -q -m gpu -s -p no:cacheprovider > gpurun_out/r2_pytest_new.txt 2>&1; grep -v "^$" gpurun_out/r2_pytest_new.txt | tail -120
echo "== bench lanes on"; D3FK_LIB=tools/libd3fk_dbg.so D3FK_SLAB_FLAT=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-cudnn --no-swap --sample-steps 50 > gpurun_out/r2_bench_lane1.txt 2>&1; tail -c 1800 gpurun_out/r2_bench_lane1.txt
echo "== bench lanes off"; D3FK_BRANCH_LANE=0 D3FK_LIB=tools/libd3fk_dbg.so D3FK_SLAB_FLAT=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-cudnn --no-swap --sample-steps 50 > gpurun_out/r2_bench_lane0.txt 2>&1; tail -c 1800 gpurun_out/r2_bench_lane0.txt
