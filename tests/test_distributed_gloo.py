"""world_size-2 gloo tests of the sharded paths (CPU): the Lloyd loop with its per-iteration
all-reduce of [sums | counts | n_changed], kernels running under the fiber emulation.
uint8 rows: the sums are exact integers in float64, so the 2-rank result must equal the
1-rank result bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case():
    rng = np.random.default_rng(21)
    n, d, k = 3001, 4, 5
    cen = rng.uniform(20, 230, (k, d))
    X = np.clip(np.rint(cen[rng.integers(k, size=n)] + rng.normal(0, 14, (n, d))), 0, 255).astype(np.uint8)
    return X, X[:k].astype(np.float64)


def _case_empty():
    """initial centres with two far-away outliers: both go empty in the first E-step and must be relocated"""
    rng = np.random.default_rng(22)
    n, d, k = 2500, 4, 6
    cen = rng.uniform(60, 120, (4, d))
    X = np.clip(np.rint(cen[rng.integers(4, size=n)] + rng.normal(0, 9, (n, d))), 0, 255).astype(np.uint8)
    init = np.concatenate([X[:4].astype(np.float64), np.full((2, d), 250.0) + np.arange(2)[:, None]])
    return X, init


def _case_many_empty():
    """20 clusters empty at once: more than the device-side candidate list (16) holds -> the fit is redone with the
    host-merged relocation"""
    rng = np.random.default_rng(23)
    n, d, k = 2200, 4, 26
    cen = rng.uniform(60, 120, (6, d))
    X = np.clip(np.rint(cen[rng.integers(6, size=n)] + rng.normal(0, 9, (n, d))), 0, 255).astype(np.uint8)
    init = np.concatenate([X[:6].astype(np.float64), np.full((20, d), 235.0) + np.arange(20)[:, None]])
    return X, init


def _worker(rank, world, port, out_dir, case="plain"):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from opticalflowclustering_b200 import kmeans as km
    from opticalflowclustering_b200.sharding import shard_range
    from tests.emu import emu_lib as E
    km._use_test_library(E.lib())
    X, init = {"plain": _case, "empty": _case_empty, "many": _case_many_empty}[case]()
    lo, hi = shard_range(len(X), rank, world)
    labels, centres, inertia, n_iter = km.lloyd(X[lo:hi], init, group=dist.group.WORLD)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), labels=labels.numpy(), centres=centres.numpy(),
             inertia=float(inertia), n_iter=int(n_iter), lo=lo, hi=hi)
    dist.destroy_process_group()


def test_sharded_lloyd_equals_single_rank(tmp_path):
    from tests.emu import emu_lib as E
    E.build()
    from opticalflowclustering_b200 import kmeans as km
    km._use_test_library(E.lib())
    X, init = _case()
    l1, c1, i1, n1 = km.lloyd(X, init)
    km._use_test_library(None)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = [np.load(tmp_path / f"r{r}.npz") for r in range(2)]
    labels = np.concatenate([g["labels"] for g in got])
    assert (labels == l1.numpy()).all()
    for g in got:
        assert (g["centres"] == c1.numpy()).all()          # bit-identical across world sizes
        assert int(g["n_iter"]) == int(n1)
        assert abs(float(g["inertia"]) - float(i1)) <= 1e-12 * float(i1)


def test_sharded_lloyd_relocates_empty_clusters_like_single_rank(tmp_path):
    """the relocation of empty clusters across ranks (global farthest rows, merged on the host)
    gives the single-rank result bit for bit"""
    from tests.emu import emu_lib as E
    E.build()
    from opticalflowclustering_b200 import kmeans as km
    from oracle import kmeans_np as K
    km._use_test_library(E.lib())
    X, init = _case_empty()
    l1, c1, i1, n1 = km.lloyd(X, init)
    km._use_test_library(None)
    ref = K.kmeans_fit(X, init)                              # the relocation really happens and matches sklearn's rule
    assert (l1.numpy() == ref[0]).all() and int(n1) == ref[3]
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path), "empty"), nprocs=2, join=True)
    got = [np.load(tmp_path / f"r{r}.npz") for r in range(2)]
    assert (np.concatenate([g["labels"] for g in got]) == l1.numpy()).all()
    for g in got:
        assert (g["centres"] == c1.numpy()).all() and int(g["n_iter"]) == int(n1)


def test_sharded_lloyd_many_empty_clusters_falls_back(tmp_path):
    from tests.emu import emu_lib as E
    E.build()
    from opticalflowclustering_b200 import kmeans as km
    km._use_test_library(E.lib())
    X, init = _case_many_empty()
    l1, c1, i1, n1 = km.lloyd(X, init)
    km._use_test_library(None)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path), "many"), nprocs=2, join=True)
    got = [np.load(tmp_path / f"r{r}.npz") for r in range(2)]
    assert (np.concatenate([g["labels"] for g in got]) == l1.numpy()).all()
    for g in got:
        assert (g["centres"] == c1.numpy()).all() and int(g["n_iter"]) == int(n1)


def test_shard_ranges_cover_with_halo():
    from opticalflowclustering_b200.sharding import frame_shards, shard_range
    for n, w in [(10, 3), (1000, 8), (7, 8), (65, 2)]:
        prev = 0
        for r in range(w):
            lo, hi = shard_range(n, r, w)
            assert lo == prev and hi >= lo
            prev = hi
        assert prev == n
    # frame shards: contiguous pair ranges; each rank loads one extra leading frame (the halo)
    shards = frame_shards(1000, 8)
    pairs = sum(s.n_pairs for s in shards)
    assert pairs == 999
    for a, b in zip(shards, shards[1:]):
        assert b.first_frame == a.first_frame + a.n_pairs     # b's first frame = a's last frame
    assert all(s.n_frames == s.n_pairs + 1 for s in shards if s.n_pairs)
