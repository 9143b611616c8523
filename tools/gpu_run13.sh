#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-r13}
echo "== pytest ops+unet"; timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py -q -m gpu -p no:cacheprovider -x > gpurun_out/${T}_pytest.txt 2>&1; tail -3 gpurun_out/${T}_pytest.txt
for spec in "4096 256 2304 0" "1024 512 4608 0" "16384 128 1152 0"; do
  echo "== timeline $spec"; D3FK_LIB=tools/libd3fk_tl.so timeout 300 python tools/timeline.py $spec 2>&1 | tail -4
done
echo "== split conv / bn"; timeout 600 python tools/split_convbn.py > gpurun_out/${T}_split.txt 2>&1; tail -48 gpurun_out/${T}_split.txt | awk '{print $1,$2,$3,$4,$5,$6,$7,$9,$10,$11}'
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-cudnn --no-swap --sample-steps 200 > gpurun_out/${T}_bench.txt 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/${T}_bench.txt
echo "== per-op"; timeout 600 python tools/profile_ops.py --repeat 20 --top 400 > gpurun_out/${T}_per_op.txt 2>&1; grep "====" gpurun_out/${T}_per_op.txt
