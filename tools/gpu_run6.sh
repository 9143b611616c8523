#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== DP test (2 GPUs)"; timeout 900 python -m pytest tests/test_gpu_parity_configs.py -q -m gpu -s -p no:cacheprovider -k "data_parallel" > gpurun_out/r6_dp_test.txt 2>&1; tail -12 gpurun_out/r6_dp_test.txt
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-swap --sample-steps 100"
echo "== bench N=2 eager"; timeout 600 $T > gpurun_out/r6_bench2_eager.txt 2>&1; grep '^{' gpurun_out/r6_bench2_eager.txt | cut -c1-900
echo "== bench N=2 graph"; D3FK_TRAIN_GRAPH_DP=1 timeout 600 $T > gpurun_out/r6_bench2_graph.txt 2>&1; grep '^{' gpurun_out/r6_bench2_graph.txt | cut -c1-900; tail -5 gpurun_out/r6_bench2_graph.txt | cut -c1-400
echo "== bench N=2 eager, NCCL_MAX_CTAS=8"; NCCL_MAX_CTAS=8 timeout 600 $T > gpurun_out/r6_bench2_cta8.txt 2>&1; grep '^{' gpurun_out/r6_bench2_cta8.txt | cut -c1-600
echo "== in-situ"; timeout 600 python -m pytest tests/test_gpu_unet.py -q -m gpu -p no:cacheprovider -k "in_situ" 2>&1 | tail -3
