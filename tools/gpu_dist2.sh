#!/bin/bash
# 2-GPU checks of the one collective on the path (NCCL all-reduce of the k-means sums) and of the sharded clip pipeline
mkdir -p gpurun_out
export PYTHONPATH=.
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/dist_kmeans_check.py > gpurun_out/dist_kmeans_${N}gpu.log 2>&1; echo "dist_kmeans rc=$?"
grep "world=" gpurun_out/dist_kmeans_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tools/clip_cluster.py --size 1080p --frames 65 > gpurun_out/clip_cluster_${N}gpu.log 2>&1; echo "clip_cluster rc=$?"
grep "world=" gpurun_out/clip_cluster_${N}gpu.log || tail -5 gpurun_out/clip_cluster_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 tools/clip_cluster.py --size 4k --frames 33 --chunk 9 > gpurun_out/clip_cluster_4k_${N}gpu.log 2>&1; echo "clip_cluster 4k rc=$?"
grep "world=" gpurun_out/clip_cluster_4k_${N}gpu.log || tail -5 gpurun_out/clip_cluster_4k_${N}gpu.log
