#!/bin/bash
# ncu --set full captures of the narrow decoder-tail kernels (one launch each); outputs under gpurun_out/
set -x
cap() {  # name kind M N K mode
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -c 1 -f -o gpurun_out/$1 \
    python tools/one_op.py $2 $3 $4 $5 $6 > gpurun_out/$1.log 2>&1
  tail -2 gpurun_out/$1.log
}
cap r01d_slab16_k288 conv 1048576 16 288 0
cap r01d_head_n3 conv 1048576 3 144 0
cap r01d_slab32_k1152 conv 262144 32 1152 0
cap r01d_stem conv 262144 64 392 0
