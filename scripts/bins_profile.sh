#!/bin/bash
# ncu --set full of the binning kernels (count, scatter) inside the bench command
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
T=${1:-r2bin}
$CMD > gpurun_out/${T}_plain.json 2> gpurun_out/${T}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${T}_plain.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:pc_bin_count_kernel -s 4 -c 1 -f -o gpurun_out/${T}_count $CMD > gpurun_out/${T}_count.log 2>&1; echo "count rc=$?"
ncu --set full --clock-control none --import-source on -k regex:pc_bin_scatter_kernel -s 4 -c 1 -f -o gpurun_out/${T}_scatter $CMD > gpurun_out/${T}_scatter.log 2>&1; echo "scatter rc=$?"
