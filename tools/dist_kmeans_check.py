"""2+ GPU check of the one collective on the path: Lloyd k-means over row-sharded uint8 vectors with a
per-iteration NCCL all-reduce of [sums | counts | n_changed] (kmeans.lloyd(group=...)).
Run: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dist_kmeans_check.py
Every rank clusters its shard; rank 0 also clusters the whole set alone; centres must be bit-identical."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from opticalflowclustering_b200 import kmeans as km
from opticalflowclustering_b200.sharding import shard_range

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = torch.Generator().manual_seed(5)
N, D, K = 1_000_000, 4, 8
cen = torch.randint(20, 230, (K, D), generator=g).double()
X = (cen[torch.randint(0, K, (N,), generator=g)] + 12 * torch.randn(N, D, generator=g, dtype=torch.float64)).round().clamp(0, 255).to(torch.uint8)
init = X[:K].double()
lo, hi = shard_range(N, rank, world)
Xs = X[lo:hi].cuda()
km.lloyd(Xs, init, group=dist.group.WORLD)                   # warm-up
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
labels, centres, inertia, n_iter = km.lloyd(Xs, init, group=dist.group.WORLD)
torch.cuda.synchronize(); dist.barrier()
dt = time.perf_counter() - t0
ok = True
if rank == 0:
    l1, c1, i1, n1 = km.lloyd(X.cuda(), init)
    ok = torch.equal(c1, centres) and int(n1) == int(n_iter) and torch.equal(l1[lo:hi], labels) and \
        abs(float(i1) - float(inertia)) <= 1e-12 * float(i1)
    print(f"world={world} N={N} D={D} k={K} n_iter={int(n_iter)} sharded {dt * 1e3:.1f} ms "
          f"({N * int(n_iter) / dt / 1e9:.2f} G rows/s)  centres bit-identical to 1 GPU: {ok}")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
