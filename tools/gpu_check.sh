#!/bin/bash
# One GPU call: parity tests, smoke, bench, then the ncu launch list of a short bench run.
# Usage (from the repo root on the GPU box): bash tools/gpu_check.sh [ncu]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
  SHORT="python bench.py --steps 2 --warmup 3 --chunk 9 --clip-frames 33 --no-cpu-baseline"
  $SHORT > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu.log 2>&1
  echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log
fi
