#!/bin/bash
# Round-end check on one GPU: smoke, the whole GPU suite, the default bench line, the reference arm, the other workloads.
mkdir -p gpurun_out
echo "== smoke"; timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/smoke.log
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
echo "== bench (default)"; timeout 600 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench_cfg2.json
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo "exit $?"; cat gpurun_out/bench_ref.json
for w in cfg1 cfg3 cfg4 cfg5; do echo "== bench $w"; timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2>> gpurun_out/bench.err; echo "exit $?"; cut -c1-700 gpurun_out/bench_$w.json; done
tail -5 gpurun_out/bench.err
