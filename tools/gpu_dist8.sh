#!/bin/bash
# N-GPU: NCCL k-means checks, sharded clip pipeline, then the bench at N GPUs
N=${1:-8}
bash tools/gpu_dist2.sh $N
export PYTHONPATH=.
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"
python -c "
import json; b=json.load(open('gpurun_out/bench_${N}gpu.json')); print(b['n_gpus'], b['value'], b['e2e']['value'], b['ms_per_step'])"
