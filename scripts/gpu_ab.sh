#!/bin/bash
# A/B several builds of the library (build/lib_*.so) on one workload: prints samples/s and roofline fraction.
# usage: bash scripts/gpu_ab.sh [workload] [extra bench args]
W=${1:-cfg2}; shift
mkdir -p gpurun_out
for so in build/lib_*.so; do
  tag=$(basename $so .so)
  FFB_LIB=$PWD/$so timeout 100 python bench.py --workload $W --steps 2 --warmup 2 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ab_${W}_$tag.log 2>&1
  python - "$tag" gpurun_out/ab_${W}_$tag.log <<'PY'
import json, sys
tag, p = sys.argv[1:3]
try:
    d = json.loads(open(p).read().strip().split("\n")[-1])
    r = d["roofline"]
    print(f"{tag:28s} {d['value']/1e6:8.3f} M/s  frac {r['frac']:.3f}  kernel {r['avg_launch_ms']:.3f} ms  steps {d['config']['dopri5_steps']}")
except Exception as e:
    print(tag, "FAILED", e, open(p).read()[-400:])
PY
done
