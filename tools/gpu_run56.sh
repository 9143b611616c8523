#!/bin/bash
# ncu --set full of the slab weight gradients of the decoder tail (the most exposed side-stream work: r55)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
bash tools/ncu_slab.sh "r02_wgslab_c64 wgrad 262144 32 576" "r02_wgslab_c16 wgrad 1048576 16 144"
for f in gpurun_out/r02_wgslab_*_details.txt; do echo "== $f"; grep -E "^  [a-z_]+<|Duration|DRAM Throughput|L2 Cache Throughput|Compute \(SM\) Throughput|Registers Per|Dynamic Shared|Theoretical Occ|Executed Ipc Active|No Eligible" $f | sed 's/  */ /g' | head -12; done
