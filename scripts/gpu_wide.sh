#!/bin/bash
# Wide-engine check on one B200: its own tests, then the whole GPU suite (the generic kernels moved to a header), then a timing
# of wide networks (scripts/time_wide.py).
mkdir -p gpurun_out
echo "== wide tests"; timeout 900 python -m pytest tests/test_gpu_wide.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_wide.log 2>&1; echo "exit $?"; tail -30 gpurun_out/pytest_wide.log
echo "== time wide"; timeout 300 python scripts/time_wide.py > gpurun_out/time_wide.json 2> gpurun_out/time_wide.err; echo "exit $?"; cat gpurun_out/time_wide.json; tail -3 gpurun_out/time_wide.err
if [ "$1" == "all" ]; then
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 400 --timeout-method=thread -p no:cacheprovider --deselect tests/test_gpu_wide.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
fi
