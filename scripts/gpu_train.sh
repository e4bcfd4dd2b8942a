#!/bin/bash
# Fused training step on one B200: its tests, then a timing against torch autograd on the same GPU (scripts/time_train.py).
mkdir -p gpurun_out
echo "== train tests"; timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_train.log 2>&1; echo "exit $?"; tail -40 gpurun_out/pytest_train.log
echo "== time train"; timeout 300 python scripts/time_train.py > gpurun_out/time_train.json 2> gpurun_out/time_train.err; echo "exit $?"; cat gpurun_out/time_train.json; tail -3 gpurun_out/time_train.err
