#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-2}
echo "== eval parity (new 96x96 case)"; timeout 600 python -m pytest tests/test_gpu_unet.py -q -m gpu -p no:cacheprovider -k "eval_forward_parity" 2>&1 | tail -2
echo "== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-cudnn --no-swap --sample-steps 200 > gpurun_out/r50_bench$N.txt 2>&1; grep -o '"value": [0-9.]*, "unit": "img/s", "n_gpus": [0-9]*[^}]*"ms_per_step": [0-9.]*' gpurun_out/r50_bench$N.txt | head -1; grep -o '"sample": {"metric": "sample_img_steps_per_s", "value": [0-9.]*' gpurun_out/r50_bench$N.txt
