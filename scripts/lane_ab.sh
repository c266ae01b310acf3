for m in 0 1 2 4 8 16 32 63; do
  for res in 256 128; do
    echo "== mask $m res $res"; SENAS_DW_LANE=$m python scripts/profile_cell.py bf16 16 3 $res 2>/dev/null | grep -E "^mode|^dw_"
  done
done
