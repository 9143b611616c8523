#!/bin/bash
export PYTHONUNBUFFERED=1
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu --no-cudnn --no-swap --no-sample"
run() { echo -n "$1: "; shift; env "$@" timeout 300 $B 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1; }
run base X=1
run max_ctas4 NCCL_MAX_CTAS=4
run max_ctas2 NCCL_MAX_CTAS=2
run min_ctas16 NCCL_MIN_CTAS=16
run ll128 NCCL_PROTO=LL128
run simple NCCL_PROTO=Simple
run nvls0 NCCL_NVLS_ENABLE=0
run eager_dp D3FK_TRAIN_GRAPH_DP=0
