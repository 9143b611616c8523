#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu --no-cudnn --no-swap --sample-steps 100 > gpurun_out/r66_bench$N.txt 2>&1
grep -o '"value": [0-9.]*, "unit": "img/s", "n_gpus": [0-9]*[^}]*"ms_per_step": [0-9.]*' gpurun_out/r66_bench$N.txt | head -1; grep -o '"sample": {"metric": "sample_img_steps_per_s", "value": [0-9.]*' gpurun_out/r66_bench$N.txt; tail -2 gpurun_out/r66_bench$N.txt | cut -c1-300
