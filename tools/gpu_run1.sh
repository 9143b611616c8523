#!/bin/bash
# first GPU pass of round 2: tests, parity evidence, bench (graph on / off), per-op profile, ncu captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1_smi.txt 2>&1
echo "== pytest new files"; timeout 900 python -m pytest tests/test_gpu_parity_configs.py -q -m gpu -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/r1_pytest_new.txt; tail -5 gpurun_out/r1_pytest_new.txt
echo "== pytest rest"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_gpu_parity_configs.py 2>&1 | tail -40 > gpurun_out/r1_pytest_rest.txt; tail -5 gpurun_out/r1_pytest_rest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.txt 2>&1; tail -3 gpurun_out/r1_smoke.txt
echo "== bf16 evidence"; timeout 600 python tools/diag_bf16_evidence.py > gpurun_out/r1_bf16_evidence.txt 2>&1; tail -20 gpurun_out/r1_bf16_evidence.txt
echo "== bench (graph)"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r1_bench_graph.txt 2>&1; tail -c 3000 gpurun_out/r1_bench_graph.txt
echo "== bench (eager)"; D3FK_TRAIN_GRAPH=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-cudnn --no-swap --no-sample > gpurun_out/r1_bench_eager.txt 2>&1; tail -c 1500 gpurun_out/r1_bench_eager.txt
echo "== bench (eager, grouped wgrad)"; D3FK_TRAIN_GRAPH=0 D3FK_WGRAD_GROUP=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-cudnn --no-swap --no-sample > gpurun_out/r1_bench_eager_group.txt 2>&1; tail -c 1500 gpurun_out/r1_bench_eager_group.txt
echo "== bench (graph, grouped wgrad)"; D3FK_WGRAD_GROUP=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-cudnn --no-swap --no-sample > gpurun_out/r1_bench_graph_group.txt 2>&1; tail -c 1500 gpurun_out/r1_bench_graph_group.txt
echo "== per-op"; timeout 600 python tools/profile_ops.py --repeat 20 --top 400 > gpurun_out/r1_per_op.txt 2>&1; tail -3 gpurun_out/r1_per_op.txt
echo "== timeline"; D3FK_TRAIN_GRAPH=0 timeout 300 python tools/step_timeline.py > gpurun_out/r1_timeline.txt 2>&1; cat gpurun_out/r1_timeline.txt | tail -15
echo "== ncu"
bash tools/ncu_slab.sh "r2_slab64_l1 conv 65536 64 576 0" "r2_tc128_l3 conv 4096 256 2304 0" "r2_slab16_k288 conv 1048576 16 288 0" "r2_stem conv 262144 64 392 0" "r2_dgrad_l3 conv 4096 256 2304 1" 2>&1 | tail -10
ls -la gpurun_out | head -60
