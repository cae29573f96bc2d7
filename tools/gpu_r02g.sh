#!/bin/bash
# 2-GPU call: two devices in one process, NCCL k-means check, bench under torchrun, NVDEC probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | head -3
python -m pytest tests -m gpu -x -q -k "two_devices or more_than_4096" 2>&1 | tail -4
python tools/nvdec_probe.py > gpurun_out/r02g_nvdec.json 2>&1; cat gpurun_out/r02g_nvdec.json
for ru in 4 2; do echo "RU=$ru"; OFC_CELLS_RU=$ru python tools/cells_bench.py k8 2>&1 | tail -2; done
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$T tools/dist_kmeans_check.py > gpurun_out/r02g_dist_kmeans_2gpu.log 2>&1; echo "dist check rc=$?"; tail -3 gpurun_out/r02g_dist_kmeans_2gpu.log
$T bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02g_bench_2gpu.json 2> gpurun_out/r02g_bench_2gpu.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r02g_bench_2gpu.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/r02g_bench_2gpu.json"))
print("value",b["value"],"e2e",b["e2e"]["value"],"affinity",b["config"].get("rank_cpu_affinity"))
for n,v in b["extras"]["dist_kmeans"].items(): print(n,"iters",v["n_iter"],"fit ms",round(v["fit_ms"],3),"Grows/s",round(v["rows_iter_per_s"]/1e9,2))
print(b["extras"]["h2d_pinned_all_ranks"])
PY
