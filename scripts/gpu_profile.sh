#!/bin/bash
# Run on the GPU box (gpurun -- 'bash scripts/gpu_profile.sh TAG'): plain bench run, then the ncu launch list of the
# same command, then one --set full capture of the dominant kernel.  Outputs land in gpurun_out/.
TAG=${1:-rX}
KREGEX=${2:-pc_query_packet_kernel}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
cat gpurun_out/${TAG}_plain.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s 3 -c 2 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
