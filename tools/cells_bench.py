"""GPU check + timing of the per-cell k-means kernels (fast shared-memory form vs the first kernel).
Scratch tool for gpurun; prints JSON lines."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowclustering_b200 import grid, kmeans as km          # noqa: E402
from opticalflowclustering_b200.pipeline import ClipPipeline      # noqa: E402
from opticalflowclustering_b200.synthetic import synthetic_clip   # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(0)
    only_k8 = len(sys.argv) > 1 and sys.argv[1] == "k8"
    for k in ((8,) if only_k8 else (3, 8, 16)):
        cells = torch.randint(0, 256, (350, 76 * 77, 4), generator=g).to(torch.uint8).to(dev)
        os.environ["OFC_CELLS_FAST"] = "1"
        t_fast, a = timed(lambda: km.lloyd_cells(cells, k, seed=1))
        os.environ["OFC_CELLS_FAST"] = "0"
        t_slow, b = timed(lambda: km.lloyd_cells(cells, k, seed=1))
        os.environ.pop("OFC_CELLS_FAST")
        same = all(torch.equal(u, v) for u, v in zip(a, b))
        print(json.dumps({"case": f"random 350x5852x4 k={k}", "fast_ms": t_fast, "first_kernel_ms": t_slow, "identical": same,
                          "mean_iters": float(a[3].float().mean())}), flush=True)
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        return
    if only_k8:
        H, W, F = 1080, 1920, 9
        clip = synthetic_clip(F, H, W, seed=3, device=dev)
        pipe = ClipPipeline(W, H, chunk_frames=F, device=dev)
        pipe.run_chunk(clip)
        viz = pipe.viz.clone()
        t, out = timed(lambda: grid.grid_kmeans_cells(viz, 8, seed=0))
        print(json.dumps({"case": "1080p viz k=8", "fused_ms_per_frame": t / (F - 1), "mean_iters": float(out["n_iter"].float().mean())}), flush=True)
        return
    for name, (H, W) in {"720p": (720, 1280), "1080p": (1080, 1920)}.items():
        F = 9
        clip = synthetic_clip(F, H, W, seed=3, device=dev)
        pipe = ClipPipeline(W, H, chunk_frames=F, device=dev)
        pipe.run_chunk(clip)
        viz = pipe.viz.clone()
        for k in (2, 8):
            t, out = timed(lambda: grid.grid_kmeans_cells(viz, k, seed=0, want_centres=True))
            # against gather + first kernel
            n = (H // 14) * (W // 25)
            ex = torch.empty((F - 1, 350, n, 4), dtype=torch.uint8, device=dev)
            from opticalflowclustering_b200 import _lib
            from opticalflowclustering_b200.flow import _ptr, _stream_ptr
            _lib.check(_lib.lib().ofc_grid_extract_cells(_ptr(viz), F - 1, H, W, 14, 25, 1, 30, 0, _ptr(ex), _stream_ptr()))
            os.environ["OFC_CELLS_FAST"] = "0"
            t_old, ref = timed(lambda: km.lloyd_cells(ex.view(-1, n, 4), k, seed=0), reps=1)
            os.environ.pop("OFC_CELLS_FAST")
            same = torch.equal(out["centres"].view(-1, k, 4), ref[1]) and torch.equal(out["n_iter"].view(-1), ref[3])
            print(json.dumps({"case": f"{name} viz, {F - 1} frames, k={k}", "fused_ms_per_frame": t / (F - 1),
                              "gather_plus_first_kernel_ms_per_frame": t_old / (F - 1), "identical": same,
                              "mean_iters": float(out["n_iter"].float().mean()), "max_iters": int(out["n_iter"].max())}), flush=True)
        del pipe
        for k in (1, 8):
            Fc = 17
            clip = synthetic_clip(33, H, W, seed=4, device=dev)
            pipe = ClipPipeline(W, H, chunk_frames=Fc, device=dev, n_clusters=k)
            t, _ = timed(lambda: pipe.run_chunk(clip[:Fc]), reps=5)
            print(json.dumps({"case": f"ClipPipeline {name} chunk {Fc} k={k}", "ms_per_step": t, "pairs_per_s": (Fc - 1) / (t / 1e3)}), flush=True)
            del pipe


if __name__ == "__main__":
    main()
