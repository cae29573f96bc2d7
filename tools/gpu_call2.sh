python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in 0 1 2; do OFC_ITER_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant $v value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),{k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"; done
bash tools/gpu_profile.sh
