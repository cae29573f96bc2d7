#!/bin/bash
# N-GPU check of the peer-memory exchange: protocol test on one device, sharded fits with both exchanges, bench legs
mkdir -p gpurun_out
export PYTHONPATH=.
N=${1:-2}
TAG=${2:-r02k}
timeout 300 python -m pytest tests/test_gpu_peer.py -x -q > gpurun_out/${TAG}_peer_test.log 2>&1; echo "peer test rc=$?"; tail -3 gpurun_out/${TAG}_peer_test.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/dist_kmeans_check.py > gpurun_out/${TAG}_dist_kmeans_${N}gpu.log 2>&1; echo "dist_kmeans rc=$?"
grep "world=" gpurun_out/${TAG}_dist_kmeans_${N}gpu.log || tail -20 gpurun_out/${TAG}_dist_kmeans_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "bench rc=$?"
tail -3 gpurun_out/${TAG}_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_${N}gpu.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"])
print(json.dumps(d["extras"].get("dist_kmeans"), indent=1))
PY
