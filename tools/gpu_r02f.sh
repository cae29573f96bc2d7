#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/r02f_gputests.log; tail -6 gpurun_out/r02f_gputests.log
for ru in 4 2; do for pk in 1 0; do
  echo "RU=$ru PACK=$pk"; OFC_CELLS_RU=$ru OFC_CELLS_PACK=$pk python tools/cells_bench.py k8 2>&1 | tail -4
done; done
python bench.py --steps 10 --warmup 3 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02f_bench.err
for c in f32_1Mx128_k1024 u8_8Mx4_k8 u8_64Mx4_k8; do python tools/fit_profile.py $c; OFC_KMEANS_STEP_FAST=0 python tools/fit_profile.py $c; done
