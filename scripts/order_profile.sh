#!/bin/bash
# ncu launch list restricted to the batch-ordering kernels (key kernel + sort passes) of the bench workload
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/${1}_plain.json 2> gpurun_out/${1}_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pc_query_key_kernel|os_pass|os_histogram|rs_' -c 120 --csv --log-file gpurun_out/${1}_order_launches.csv $CMD > gpurun_out/${1}_ncu.log 2>&1
echo rc=$?
