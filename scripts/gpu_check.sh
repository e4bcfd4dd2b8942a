#!/bin/bash
# One GPU-box round: smoke, GPU parity tests, a short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
echo "== smoke (engine: ${FFB_ENGINE:-tc})" ; timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke exit $rc" | tee -a gpurun_out/smoke.log
tail -8 gpurun_out/smoke.log
if [ $rc -ne 0 ] && [ -z "$KEEP_GOING" ]; then echo "smoke failed: stopping"; exit 1; fi
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps ${BENCH_STEPS:-3} --warmup ${BENCH_WARMUP:-3} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" | tee -a gpurun_out/bench.err
cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
