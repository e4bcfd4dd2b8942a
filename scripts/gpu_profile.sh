#!/bin/bash
# ncu evidence for one workload: launch list (time per launch) + one full capture of the top kernel.
# usage: bash scripts/gpu_profile.sh <workload> <batch> <kernel-regex> <tag>
W=${1:-cfg2}; B=${2:-151552}; K=${3:-k_dopri5}; TAG=${4:-r01_${W}}
mkdir -p gpurun_out
CMD="python bench.py --workload $W --steps 1 --warmup 1 --batch $B --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${TAG}_plain.log; exit 1; }
cat gpurun_out/${TAG}_plain.log | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 2 -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/
