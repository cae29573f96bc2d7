#!/bin/bash
python -c "from opticalflowclustering_b200 import _build; _build.build()"
mkdir -p gpurun_out
export OFC_ITER_TMEM=${TM:-513}
timeout 300 python -m pytest tests/test_gpu_flow.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_tmem.log; echo "pytest rc=${PIPESTATUS[0]}" >> gpurun_out/pytest_tmem.log
tail -12 gpurun_out/pytest_tmem.log
summ() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],3),{k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"; }
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | summ "tmem"
