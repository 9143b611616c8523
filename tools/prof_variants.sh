for v in "" "D3FK_A_CA=1" "D3FK_A_CA=1 D3FK_OCC=1" "D3FK_OCC=1" "D3FK_SPLIT_TILES=147"; do
  echo "=== variant: $v"
  env $v timeout 300 python tools/profile_ops.py --top 0 2>&1 | grep -E "^====|OP_CONV |OP_WGRAD " | grep -v "^     "
done
