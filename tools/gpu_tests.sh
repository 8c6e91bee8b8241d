#!/bin/bash
# the GPU test suite without -x, with durations: gpurun_out/<tag>_pytest_gpu.log
tag=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=12 > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${tag}_pytest_gpu.log | tail -40
