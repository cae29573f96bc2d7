#!/bin/bash
python -c "from opticalflowclustering_b200 import _build; _build.build()"
# A/B of the flow-iteration kernels: parity tests, then bench per variant / strip height, then ncu of the strip kernel.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; echo "pytest rc=${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
summ() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],3),{k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"; }
for v in 0 1; do OFC_ITER_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | summ "variant=$v"; done
for sh in 90 108 120; do OFC_STRIP_H=$sh python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | summ "strip_h=$sh"; done
python tools/profile_step.py > gpurun_out/profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:flow_iter_strip -s 33 -c 1 -o gpurun_out/prof_flow_iter_strip -f python tools/profile_step.py > gpurun_out/ncu_iter.log 2>&1; echo "iter rc=$?"
