"""BASELINE configs[2]/[3] end to end on N GPUs of one box: a synthetic clip is cut into contiguous frame-pair
ranges (one-frame halo, sharding.frame_shards), every rank runs flow -> visualisation -> grid -> k = 1 hues on
its range (no collective), then ALL ranks cluster the clip's per-frame hue vectors (350 uint8 features per
pair, the rows of the reference's OutCSV) with one global k-means whose only exchange is the per-iteration
NCCL all-reduce of [sums | counts | n_changed] (kmeans.lloyd(group=...)).

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/clip_cluster.py \
      [--size 1080p|4k|720p] [--frames 65] [--levels 3] [--k 8]

Rank 0 also runs the whole clip alone and checks: hue rows identical (sharding changes nothing), cluster centres
bit-identical (uint8 sums are exact integers in float64, so the all-reduce order does not matter).
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from opticalflowclustering_b200 import kmeans as km
from opticalflowclustering_b200.pipeline import ClipPipeline
from opticalflowclustering_b200.sharding import frame_shards
from opticalflowclustering_b200.synthetic import synthetic_clip

SIZES = {"720p": (1280, 720), "1080p": (1920, 1080), "4k": (3840, 2160)}


def hue_rows(clip, W, H, levels, chunk):
    pipe = ClipPipeline(W, H, chunk_frames=min(chunk, clip.shape[0]), levels=levels)
    out = pipe.process_clip(clip)
    return out["km_hue"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="1080p", choices=sorted(SIZES))
    ap.add_argument("--frames", type=int, default=65)
    ap.add_argument("--levels", type=int, default=None)
    ap.add_argument("--k", type=int, default=8)
    ap.add_argument("--chunk", type=int, default=17)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H = SIZES[args.size]
    levels = args.levels if args.levels is not None else (5 if args.size == "4k" else 3)
    clip = synthetic_clip(args.frames, H, W, seed=0, device="cuda")            # every rank builds the same clip
    sh = frame_shards(args.frames, world)[rank]
    mine = clip[sh.first_frame: sh.first_frame + sh.n_frames]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    rows = hue_rows(mine, W, H, levels, args.chunk).cuda()                     # [n_pairs, 350] uint8
    torch.cuda.synchronize()
    t_flow = time.perf_counter() - t0
    # identical initial centres on every rank: the first k hue rows of the clip (rank 0's), broadcast
    # (a rank may hold fewer than k rows: every rank offers its first k, the first k in rank order are taken)
    head = torch.zeros((args.k, rows.shape[1]), dtype=torch.float64, device="cuda")
    n_head = min(args.k, rows.shape[0])
    head[:n_head] = rows[:n_head].double()
    if world > 1:
        heads = torch.empty((world,) + tuple(head.shape), dtype=torch.float64, device="cuda")
        counts = torch.empty(world, dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(heads, head)
        dist.all_gather_into_tensor(counts, torch.tensor([n_head], dtype=torch.int64, device="cuda"))
        init = torch.cat([heads[r, :int(counts[r])] for r in range(world)])[:args.k].contiguous()
    else:
        init = head
    group = dist.group.WORLD if world > 1 else None
    t0 = time.perf_counter()
    labels, centres, inertia, n_iter = km.lloyd(rows, init, group=group)
    torch.cuda.synchronize()
    t_km = time.perf_counter() - t0
    ok = True
    if rank == 0:
        all_rows = hue_rows(clip, W, H, levels, args.chunk).cuda()
        same_rows = torch.equal(all_rows[: sh.n_pairs], rows)
        l1, c1, i1, n1 = km.lloyd(all_rows, init)
        ok = same_rows and torch.equal(c1, centres) and int(n1) == int(n_iter) and torch.equal(l1[: sh.n_pairs], labels)
        print(f"world={world} {args.size} levels={levels} frames={args.frames}: rank 0 did {sh.n_pairs} pairs in {t_flow * 1e3:.1f} ms "
              f"({sh.n_pairs / t_flow:.0f} pairs/s/GPU); global k-means k={args.k} over {args.frames - 1} x {rows.shape[1]} hue rows: "
              f"{int(n_iter)} iterations, {t_km * 1e3:.1f} ms; sharded hue rows identical: {same_rows}; centres bit-identical to 1 GPU: "
              f"{torch.equal(c1, centres)}; ok: {ok}")
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
