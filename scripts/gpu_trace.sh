#!/bin/bash
# GPU round for the staged Hutch++ / XTrace path: its tests first, then the whole GPU suite, smoke and a bench line.
mkdir -p gpurun_out
echo "== trace estimator tests"; timeout 600 python -m pytest tests/test_trace_estimators.py -m gpu -q --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_trace.log 2>&1; echo "exit $?" | tee -a gpurun_out/pytest_trace.log
tail -40 gpurun_out/pytest_trace.log
echo "== smoke"; timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log; tail -5 gpurun_out/smoke.log
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" | tee -a gpurun_out/bench.err
cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
