#!/bin/bash
# ncu --set full of the ordering-pass kernels (key kernel, 16-item sort pass) inside the bench command
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
T=${1:-r2o}
$CMD > gpurun_out/${T}_plain.json 2> gpurun_out/${T}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${T}_plain.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:pc_query_key_kernel -s 4 -c 1 -f -o gpurun_out/${T}_key $CMD > gpurun_out/${T}_key.log 2>&1; echo "key rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'os_pass<unsigned int, 16>|os_passIjLi16' -s 12 -c 2 -f -o gpurun_out/${T}_pass $CMD > gpurun_out/${T}_pass.log 2>&1; echo "pass rc=$?"
