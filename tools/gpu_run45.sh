#!/bin/bash
export PYTHONUNBUFFERED=1 D3FK_LIB=tools/libd3fk_dbg.so D3FK_FORK_WGRAD=0
for cfg in "4 3" "2 3" "1 3" "1 4" "2 4" "4 4"; do
  set -- $cfg
  echo "== SMAX=$1 STAGES=$2"; D3FK_WGS_SMAX=$1 D3FK_WGS_STAGES=$2 D3FK_VERBOSE=0 timeout 300 python tools/wgrad_slab_variants.py 2>&1 | grep "wgrad\|total"
done
