#!/bin/bash
# GPU round for the device-side dopri5 controller: parity tests, then cfg2 / cfg3 / cfg1 with both controllers.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 180 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
for w in cfg2 cfg1 cfg3; do
  for c in host device; do
    FFB_CONTROLLER=$c timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ctl_${w}_${c}.json 2> gpurun_out/ctl_${w}_${c}.err
    echo "$w $c exit $?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ctl_${w}_${c}.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("  value %.4g  ms/step %.2f  e2e %.4g  kernel avg %.4f ms x %d  frac %.3f share %.3f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["avg_launch_ms"], r["launches"], r["frac"], r["share_of_step"], d["gpu_launches"]))
except Exception as e:
    print("  parse failed", e); print(open("gpurun_out/ctl_${w}_${c}.err").read()[-1500:])
PY
  done
done
