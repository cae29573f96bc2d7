"""One whole Lloyd fit of a bench `dist_kmeans` case on one GPU (for an ncu launch list / profile).
    python tools/fit_profile.py f32_1Mx128_k1024 | u8_8Mx4_k8 | u8_64Mx4_k8"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from opticalflowclustering_b200 import kmeans as km

name = sys.argv[1] if len(sys.argv) > 1 else "f32_1Mx128_k1024"
N, D, K, is_u8 = {"u8_8Mx4_k8": (8_000_000, 4, 8, True), "u8_64Mx4_k8": (64_000_000, 4, 8, True),
                  "f32_1Mx128_k1024": (1_000_000, 128, 1024, False)}[name]
dev = torch.device("cuda")
gd = torch.Generator(device=dev).manual_seed(11)
cen = torch.rand((K, D), device=dev, generator=gd) * (200 if is_u8 else 8) + (25 if is_u8 else 0)
rows, init = [], None
blk = 4_000_000 if is_u8 else 250_000
for s0 in range(0, N, blk):
    nb = min(blk, N - s0)
    lab = torch.randint(0, K, (nb,), device=dev, generator=gd)
    x = cen[lab] + (12 if is_u8 else 1) * torch.randn((nb, D), device=dev, generator=gd)
    if s0 == 0:
        init = (x[:K].round().clamp(0, 255) if is_u8 else x[:K].float()).double()
    rows.append(x.round().clamp(0, 255).to(torch.uint8) if is_u8 else x.float())
X = torch.cat(rows).contiguous()
del rows
max_iter = int(os.environ.get("OFC_FIT_ITERS", "300"))
torch.cuda.synchronize()
t0 = time.perf_counter()
labels, centres, inertia, n_iter = km.lloyd(X, init, max_iter=max_iter)
torch.cuda.synchronize()
print(name, "n_iter", int(n_iter), "ms", 1e3 * (time.perf_counter() - t0), "inertia", float(inertia))
