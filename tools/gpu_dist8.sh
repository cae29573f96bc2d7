#!/bin/bash
# N-GPU: k-means exchange checks (NCCL collectives and the NVLink peer-memory kernel), sharded clip pipeline, bench at N GPUs
N=${1:-8}
TAG=${2:-r02x}
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/dist_kmeans_check.py > gpurun_out/${TAG}_dist_kmeans_${N}gpu.log 2>&1; echo "dist_kmeans rc=$?"
grep "world=" gpurun_out/${TAG}_dist_kmeans_${N}gpu.log || tail -20 gpurun_out/${TAG}_dist_kmeans_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tools/clip_cluster.py --size 1080p --frames 65 > gpurun_out/${TAG}_clip_cluster_${N}gpu.log 2>&1; echo "clip_cluster rc=$?"
grep "world=" gpurun_out/${TAG}_clip_cluster_${N}gpu.log || tail -5 gpurun_out/${TAG}_clip_cluster_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 tools/clip_cluster.py --size 4k --frames 33 --chunk 9 > gpurun_out/${TAG}_clip_cluster_4k_${N}gpu.log 2>&1; echo "clip_cluster 4k rc=$?"
grep "world=" gpurun_out/${TAG}_clip_cluster_4k_${N}gpu.log || tail -5 gpurun_out/${TAG}_clip_cluster_4k_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_${N}gpu.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"])
print(json.dumps(d["extras"].get("dist_kmeans"), indent=1))
print(json.dumps(d["extras"].get("h2d_pinned_all_ranks")))
PY
