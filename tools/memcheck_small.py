"""Tiny shapes through the kernels added late in round 1, for `compute-sanitizer --tool memcheck`:
tensor-core k-means (float32 / uint8, ring and centre-resident forms, ragged n / d / k), CSR M-step, fused uint8
step, medium E-step, cross-shape Farneback flags (Gaussian window, initial flow, run-time-radius box).
(compute-sanitizer is closed on this GPU pool, so in round 1 the driver only ran plain -- every result checked against the
CUDA-core kernels / oracle-pinned paths, no CUDA fault on the ragged shapes.)"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("tc_check", os.path.join(ROOT, "tools", "tc_check.py"))
tc = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tc)

ok = True
for shape in [(300, 36, 2), (1000, 64, 64), (777, 100, 40), (900, 128, 300), (513, 256, 130)]:
    ok = tc.run(*shape) and ok
for shape in [(500, 36, 8), (700, 352, 8), (600, 128, 256)]:
    ok = tc.run_u8(*shape) and ok

from opticalflowclustering_b200 import kmeans as km
rng = np.random.default_rng(0)
for dt, (n, d, k, B) in [(np.uint8, (3000, 4, 8, 2)), (np.uint8, (1000, 350, 8, 1)), (np.float32, (800, 70, 9, 2)), (np.uint8, (700, 33, 12, 1))]:
    X = (rng.integers(0, 256, (B, n, d)) if dt == np.uint8 else rng.normal(0, 3, (B, n, d))).astype(dt)
    km.lloyd(torch.from_numpy(X).cuda(), X[:, :k].astype(np.float64), max_iter=4)

from opticalflowclustering_b200.flow import calc_optical_flow_farneback
from opticalflowclustering_b200.synthetic import synthetic_clip
clip = synthetic_clip(3, 77, 131, seed=4).numpy()
g = clip[..., 1].copy()
f0 = calc_optical_flow_farneback(g[0], g[1], None, 0.5, 3, 15, 3, 5, 1.2, 0)
for flags, ws in ((256, 15), (4, 15), (260, 9), (0, 10), (0, 17), (0, 33)):
    calc_optical_flow_farneback(g[1], g[2], f0.copy(), 0.5, 2, ws, 2, 5, 1.2, flags)
torch.cuda.synchronize()
print("MEMCHECK_DRIVER", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
