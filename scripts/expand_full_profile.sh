#!/bin/bash
# ncu --set full captures of the sample-stream kernels of pc_expand_batch (one gpurun call; outputs in gpurun_out/)
CMD="python scripts/expand_profile.py --once 10000000"
T=${1:-r2x}
$CMD > gpurun_out/${T}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${T}_plain.log; exit 1; }
for K in pc_sample_emit_kernel pc_sample_mask_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/${T}_$K $CMD > gpurun_out/${T}_$K.log 2>&1
  echo "$K rc=$?"
done
