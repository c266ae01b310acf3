#!/bin/bash
# measurement artefacts: DRAM traffic of one whole search step, ncu of the tcgen05 kernels, timeline, per-cell graph times
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python scripts/ncu_step.py bf16 16 > $O/r2f_ncu_step_plain.log 2>&1; echo "plain_rc=$?"
timeout 1200 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
  --clock-control none --csv --log-file $O/r2f_step_traffic.csv python scripts/ncu_step.py bf16 16 > $O/r2f_ncu_step.log 2>&1; echo "traffic_rc=$?"
gzip -f $O/r2f_step_traffic.csv
timeout 600 ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section WarpStateStats \
  --section LaunchStats --section Occupancy --clock-control none -k regex:'conv_tc' --launch-skip 30 -c 30 -f -o $O/r2f_tc \
  python scripts/profile_cell.py bf16 16 2 256 > $O/r2f_ncu_tc.log 2>&1; echo "tc_rc=$?"
ncu -i $O/r2f_tc.ncu-rep --page raw --csv > $O/r2f_tc_raw.csv 2>/dev/null; gzip -f $O/r2f_tc_raw.csv
python scripts/timeline_step.py bf16 16 > $O/r2f_timeline.log 2>&1; echo "timeline_rc=$?"
python scripts/profile_cells.py > $O/r2f_cells.log 2>&1; echo "cells_rc=$?"
ls -la $O | grep r2f
echo done
