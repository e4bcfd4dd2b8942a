#!/bin/bash
mkdir -p gpurun_out
P=build/tc_probe
{
for mode in 1 3 5 13; do timeout 60 $P 128 128 2048 128 3 $mode; done
} > gpurun_out/tc_probe.log 2>&1
grep rms gpurun_out/tc_probe.log
