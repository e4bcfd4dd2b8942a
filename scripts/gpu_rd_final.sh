#!/bin/bash
# Dual-tile engine vs the single-tile engine on every workload, then the GPU test suite.
mkdir -p gpurun_out
for W in cfg2 cfg1 cfg4 cfg5; do
  for E in rd rr; do
    if [ $E == rr ]; then export FFB_ENGINE=rr; else unset FFB_ENGINE; fi
    B=""; [ $W == cfg4 ] && B="--batch 606208"; [ $W == cfg5 ] && B="--batch 1212416"
    timeout 150 python bench.py --workload $W $B --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/fin_${W}_$E.log 2>&1
    python - $W $E gpurun_out/fin_${W}_$E.log <<'PY'
import json, sys
w, e, p = sys.argv[1:4]
try:
    d = json.loads(open(p).read().strip().split("\n")[-1]); r = d["roofline"]
    print(f"{w} {e}: {d['value']/1e6:8.3f} M/s  frac {r['frac']:.3f}  kernel {r['avg_launch_ms']:.3f} ms  ms/step {d['ms_per_step']:.2f}")
except Exception as ex:
    print(w, e, "FAILED", ex, open(p).read()[-600:])
PY
  done
done
unset FFB_ENGINE
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 90 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
