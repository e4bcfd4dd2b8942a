#!/bin/bash
# A/B the build/lib_*.so variants on cfg2 (parity check of each through smoke first), then the timeline of build/lib_*trace.so
mkdir -p gpurun_out
for so in build/lib_*.so; do
  case $so in *trace*) continue;; esac
  echo "== smoke $so"; FFB_LIB=$PWD/$so timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | head -3
done
bash scripts/gpu_ab.sh cfg2
for so in build/lib_*trace*.so; do
  [ -f $so ] || continue
  FFB_LIB=$PWD/$so timeout 120 python scripts/trace_rd.py dopri5 0 4000 > gpurun_out/trace_$(basename $so .so).txt 2>&1; tail -1 gpurun_out/trace_$(basename $so .so).txt
done
