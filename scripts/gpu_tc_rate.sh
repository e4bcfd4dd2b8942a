#!/bin/bash
mkdir -p gpurun_out
{
for N in 128 64 32; do for mode in 0 1 2 3 4; do timeout 30 build/tc_rate $N $mode 96; done; done
} > gpurun_out/tc_rate.log 2>&1
cat gpurun_out/tc_rate.log
