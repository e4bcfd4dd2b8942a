#!/bin/bash
mkdir -p gpurun_out
P=build/tc_probe
{
for mode in 0 1 2 3; do timeout 60 $P 128 128 2048 128 3 $mode; done
timeout 60 $P 16 128 256 128 3 3
timeout 60 $P 128 24 2048 128 3 1
} > gpurun_out/tc_probe.log 2>&1
cat gpurun_out/tc_probe.log
