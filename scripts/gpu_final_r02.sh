#!/bin/bash
# Round-2 closing run on one B200: smoke, the whole GPU suite, the default bench line and one line per other workload.
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 400 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bench default"; timeout 400 python bench.py > gpurun_out/r02_bench_cfg2.json 2> gpurun_out/r02_bench_cfg2.err; echo "exit $?"
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_cfg2_reference_arm.json 2>/dev/null; echo "exit $?"
for W in cfg1 cfg3 cfg4 cfg5; do
  timeout 400 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/r02_bench_$W.json 2> gpurun_out/r02_bench_$W.err; echo "$W exit $?"
done
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/r02_bench_*.json")):
    try:
        d = json.loads(open(p).read().strip().split("\n")[-1]); r = d.get("roofline") or {}
        print(f"{p.split('/')[-1]:40s} {d['value']/1e6:9.3f} M/s e2e {d['e2e']['value']/1e6:9.3f} frac {r.get('frac', 0):.3f} tf32/3 {r.get('frac_of_tf32_gemm_over_3', 0):.3f} "
              f"cpu {d.get('cpu_baseline', {}).get('value', 0)/1e3:.1f}k clocks {d.get('clocks', {}).get('sm_mhz')}")
    except Exception as e:
        print(p, "FAILED", e)
PY
