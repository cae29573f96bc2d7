"""How the path is partitioned across the GPUs of one box.

Flow, visualisation, grid and per-cell k-means need no exchange: frame pairs are independent
given both frames, so a clip of T frames (T-1 pairs) is cut into contiguous pair ranges, one
per rank, and each rank loads one extra leading frame -- its first frame is the previous
rank's last (a one-frame halo, the analogue of ``prev_gray`` in
computeOpticalFlowModule.py:16,34).  Global k-means over row-sharded vectors is the one step
with a collective (kmeans.lloyd(group=...)).
"""
from __future__ import annotations

from dataclasses import dataclass


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) of n items for `rank` of `world` (first n % world ranks get one more)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


@dataclass(frozen=True)
class FrameShard:
    rank: int
    first_frame: int     # index of the first frame this rank loads (halo included)
    n_pairs: int         # pairs (first_frame+i, first_frame+i+1), i in [0, n_pairs)

    @property
    def n_frames(self) -> int:
        return self.n_pairs + 1 if self.n_pairs else 0


def frame_shards(n_frames: int, world: int):
    """Pair ranges of a clip for every rank; pair p = (frame p, frame p+1)."""
    out = []
    for r in range(world):
        lo, hi = shard_range(max(n_frames - 1, 0), r, world)
        out.append(FrameShard(r, lo, hi - lo))
    return out
