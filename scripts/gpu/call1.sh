#!/bin/bash
# round 2, call 1: bf16-z A/B (head cell + bench), ncu --set full of the head cell
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
for z in 0 1; do
  SENAS_Z_BF16=$z python scripts/profile_cell.py bf16 16 3 256 > $O/r2b_cell256_z$z.log 2>&1
  SENAS_Z_BF16=$z python scripts/profile_cell.py bf16 16 3 128 > $O/r2b_cell128_z$z.log 2>&1
  SENAS_Z_BF16=$z timeout 600 python bench.py --no-cpu --no-ref-gpu --no-fp32-line > $O/r2b_bench_z$z.json 2> $O/r2b_bench_z$z.err
done
timeout 900 ncu --set full --clock-control none -k regex:'dw_|pw_|node_|gather|conv_|reduce|lin_bwd|up8|adapter|pack|cast|bn_|pool8' \
  --launch-skip 267 -c 267 -f -o $O/r2b_cell256 python scripts/profile_cell.py bf16 16 2 256 > $O/r2b_ncu_cell256.log 2>&1
ncu -i $O/r2b_cell256.ncu-rep --page raw --csv > $O/r2b_cell256_raw.csv 2>/dev/null
ls -la $O/r2b_cell256* 
gzip -f $O/r2b_cell256_raw.csv
[ $(stat -c %s $O/r2b_cell256.ncu-rep) -gt 40000000 ] && rm $O/r2b_cell256.ncu-rep
echo done
