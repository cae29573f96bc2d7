#!/bin/bash
python -c "from opticalflowclustering_b200 import _build; _build.build()"
# GPU call: parity tests only (all -m gpu), log to gpurun_out/pytest_gpu.log
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; echo "pytest rc=${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
