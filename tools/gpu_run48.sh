#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 120 python tools/one_op.py conv 262144 64 256 2 | tail -1
bash tools/ncu_slab.sh "r02_stem_s2d conv 262144 64 256 2" "r02_wgrad_tma_l2 wgrad 16384 128 1152" "r02_conv_l3_fused conv 4096 256 2304 0" "r02_dec1_conv1_tma conv 16384 128 3456 0" "r02_slab_l1 conv 65536 64 576 0" "r02_slab_t16 conv 1048576 16 288 0"
for f in gpurun_out/r02_*_details.txt; do echo "== $f"; grep -E "^  [a-z_]+<|Duration|DRAM Throughput|L2 Cache Throughput|Compute \(SM\) Throughput|Registers Per|Dynamic Shared|Theoretical Occ|Executed Ipc Active|No Eligible" $f | sed 's/  */ /g' | head -12; done
