#!/bin/bash
# Round-2 second half on one B200: smoke, the whole GPU suite, bench lines, and ncu evidence for the two new kernels
# (wide-engine fixed-grid solve, fused training step).  Each ncu run follows a plain run of the same command that exited 0.
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q --timeout 400 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
echo "== bench default"; timeout 400 python bench.py > gpurun_out/r02b_bench_cfg2.json 2> gpurun_out/r02b_bench_cfg2.err; echo "exit $?"; tail -c 600 gpurun_out/r02b_bench_cfg2.json
for W in cfg3 cfg4; do
  timeout 400 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/r02b_bench_$W.json 2> gpurun_out/r02b_bench_$W.err; echo "$W exit $?"
done
echo "== wide / train timing"
timeout 300 python scripts/time_wide.py > gpurun_out/r02_wide_engine_timing.json 2> gpurun_out/time_wide.err; echo "exit $?"
timeout 300 python scripts/time_train.py > gpurun_out/r02_train_step_timing.json 2> gpurun_out/time_train.err; echo "exit $?"; cat gpurun_out/r02_train_step_timing.json
echo "== ncu wide"
CMDW="python scripts/time_wide.py 37888"
ncu --set full --clock-control none --import-source on -k regex:k_fixed -s 6 -c 1 -o gpurun_out/r02_wide_full $CMDW > gpurun_out/ncu_wide.log 2>&1; echo "ncu wide exit $?"
echo "== ncu train"
CMDT="python scripts/time_train.py"
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_train_launches.csv $CMDT > gpurun_out/ncu_train1.log 2>&1; echo "ncu train launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:k_train_fwdbwd -s 30 -c 1 -o gpurun_out/r02_train_full $CMDT > gpurun_out/ncu_train2.log 2>&1; echo "ncu train full exit $?"
ls -la gpurun_out | head -40
