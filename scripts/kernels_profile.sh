#!/bin/bash
# ncu --set full captures of the kernels beside the bench's dominant one (one gpurun call; outputs in gpurun_out/)
CMD="python scripts/kernels_once.py"
T=${1:-r2k}
$CMD > gpurun_out/${T}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${T}_plain.log; exit 1; }
for K in pc_clearance_kernel pc_range_coop_kernel pc_query_coop_kernel os_pass pc_query_key_kernel pc_tree_fit_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/${T}_$K $CMD > gpurun_out/${T}_$K.log 2>&1
  echo "$K rc=$?"
done
