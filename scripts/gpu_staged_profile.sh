#!/bin/bash
# GPU round for the staged Hutch++ / XTrace path: tests, timing next to the fused exact / Hutchinson solves, ncu evidence.
mkdir -p gpurun_out
echo "== trace estimator tests"; timeout 400 python -m pytest tests/test_trace_estimators.py -m gpu -q --timeout 120 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -8
echo "== timing"; timeout 300 python scripts/time_staged.py 1000000 2>&1 | tail -8
cp gpurun_out/staged_timing.json gpurun_out/r01_staged_timing.json
echo "== A/B: thread-per-sample estimator kernel"; FFB_TRACE_KERNEL=thread FFB_STAGED_ONLY=hutchpp_r1_m1 timeout 300 python scripts/time_staged.py 1000000 2>&1 | tail -2
cp gpurun_out/r01_staged_timing.json gpurun_out/staged_timing.json
echo "== ncu launch list (hutchpp r1 m1 only, 262144 rows)"
FFB_STAGED_ONLY=hutchpp_r1_m1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_staged_launches.csv python scripts/time_staged.py 262144 > gpurun_out/staged_ncu1.log 2>&1; echo "exit $?"
echo "== ncu full: k_trace_estimate, k_field_eval_rrt"
FFB_STAGED_ONLY=hutchpp_r1_m1 timeout 600 ncu --set full --clock-control none -k regex:k_trace_coop -s 4 -c 1 -o /tmp/trace_full python scripts/time_staged.py 262144 > gpurun_out/staged_ncu2.log 2>&1; echo "exit $?"
python scripts/ncu_summary.py /tmp/trace_full.ncu-rep gpurun_out/r01_trace_estimate_ncu_summary.txt 262144 "k_trace_coop<16,1>: Hutch++ r=1 m=1, D=16, 262144 rows, one launch" | tail -25
FFB_STAGED_ONLY=hutchpp_r1_m1 timeout 600 ncu --set full --clock-control none -k regex:k_field_eval_rrt -s 4 -c 1 -o /tmp/feval_full python scripts/time_staged.py 262144 > gpurun_out/staged_ncu3.log 2>&1; echo "exit $?"
python scripts/ncu_summary.py /tmp/feval_full.ncu-rep gpurun_out/r01_field_eval_jac_ncu_summary.txt 262144 "k_field_eval_rrt with Jacobian output: 16-D score model 4x128, 262144 rows, one launch" | tail -25
ls -la gpurun_out | tail -8
