#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== DP test (2 GPUs)"; timeout 600 python -m pytest tests/test_gpu_parity_configs.py -q -m gpu -p no:cacheprovider -k two_gpus > gpurun_out/r37_dp_test.txt 2>&1; tail -3 gpurun_out/r37_dp_test.txt
echo "== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --no-cudnn --no-swap --sample-steps 200 > gpurun_out/r37_bench2.txt 2>&1; grep -o '"value": [0-9.]*, "unit": "img/s", "n_gpus": 2[^}]*"ms_per_step": [0-9.]*' gpurun_out/r37_bench2.txt | head -2; tail -c 600 gpurun_out/r37_bench2.txt
