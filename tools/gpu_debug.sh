#!/bin/bash
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 python -m pytest tests/test_gpu_flow.py -m gpu -x -q > gpurun_out/debug_pytest.log 2>&1; echo "rc=$?"
grep -n "Error\|error\|assert" gpurun_out/debug_pytest.log | head -20
tail -30 gpurun_out/debug_pytest.log
