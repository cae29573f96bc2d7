#!/bin/bash
# tensor-core k-means tests + timing, fused uint8 step / k-means extras of the bench
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 600 python -m pytest tests/test_gpu_kmeans_tc.py tests/test_gpu_kmeans_cosine.py -x -q > gpurun_out/pytest_tc.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_tc.log
timeout 300 python tools/tc_check.py time > gpurun_out/tc_check.log 2>&1; echo "tc_check rc=$?"
grep -A3 "^n=1000000\|^n=200000" gpurun_out/tc_check.log | grep "tc_assign\|^n=" | cut -c1-200; tail -1 gpurun_out/tc_check.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; echo "bench rc=$?"
python -c "
import json; b=json.load(open('gpurun_out/bench_fused.json')); print(b['value']); [print(k, v) for k,v in b['extras'].items() if k.startswith('kmeans_iter')]"
