#!/bin/bash
# First GPU round of the dual-tile engine: smoke, engine cross-checks, a short bench for both engines.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke exit $rc" | tee -a gpurun_out/smoke.log
tail -8 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed"; fi
echo "== bench rd"; timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_rd.log 2> gpurun_out/bench_rd.err; echo "bench exit $?"
cat gpurun_out/bench_rd.log; tail -5 gpurun_out/bench_rd.err
echo "== bench rr"; FFB_ENGINE=rr timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_rr.log 2> gpurun_out/bench_rr.err; echo "bench exit $?"
cat gpurun_out/bench_rr.log; tail -5 gpurun_out/bench_rr.err
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 180 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
