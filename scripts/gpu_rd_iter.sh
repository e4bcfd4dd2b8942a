#!/bin/bash
# Iteration round for the dual-tile engine: smoke, GPU tests, bench of the main library and of the build/lib_*.so variants, trace.
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
echo "== bench main"; timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_main.log 2> gpurun_out/bench_main.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_main.log").read().strip().split("\n")[-1]); r = d["roofline"]
    print(f"main: {d['value']/1e6:.3f} M/s frac {r['frac']:.3f} kernel {r['avg_launch_ms']:.3f} ms steps {d['config']['dopri5_steps']} e2e {d['e2e']['value']/1e6:.3f}")
except Exception as e:
    print("bench main FAILED", e); print(open("gpurun_out/bench_main.err").read()[-1500:])
PY
for W in ${AB_WORKLOADS:-cfg2}; do bash scripts/gpu_ab.sh $W; done
if [ -f build/lib_w2_trace.so ]; then FFB_LIB=$PWD/build/lib_w2_trace.so timeout 200 python scripts/trace_rd.py dopri5 0 4000 > gpurun_out/trace_rd.txt 2>&1; tail -2 gpurun_out/trace_rd.txt; fi
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 60 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
