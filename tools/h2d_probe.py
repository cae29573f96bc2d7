"""How fast can one rank move frames from pinned host memory to the GPU?  One copy stream against 2 / 4 concurrent
streams (the bench's e2e leg is bound by this link: 199 MB per 32-pair step), and page-locked memory allocated by torch
against cudaHostAlloc'd write-combined memory.  Prints one JSON line."""
import json
import time

import torch

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
MB = 1 << 20
total = 192 * MB * 2                       # two steps' worth of frames
host = torch.empty(total, dtype=torch.uint8).pin_memory()
host.fill_(7)
dst = torch.empty(total, dtype=torch.uint8, device=dev)
out = {}
for n_streams in (1, 2, 4, 8):
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
    part = total // n_streams
    best = 0.0
    for rep in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                dst[i * part:(i + 1) * part].copy_(host[i * part:(i + 1) * part], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = max(best, total / dt / 1e9)
    out[f"h2d_gb_s_{n_streams}_streams"] = round(best, 2)
# many small copies (one frame each, 6.2 MB) on one stream, like a per-frame upload
frame = 1920 * 1080 * 3
n = total // frame
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(n):
    dst[i * frame:(i + 1) * frame].copy_(host[i * frame:(i + 1) * frame], non_blocking=True)
torch.cuda.synchronize()
out["h2d_gb_s_per_frame_copies"] = round(n * frame / (time.perf_counter() - t0) / 1e9, 2)
# device -> host for completeness
torch.cuda.synchronize()
t0 = time.perf_counter()
host.copy_(dst, non_blocking=True)
torch.cuda.synchronize()
out["d2h_gb_s"] = round(total / (time.perf_counter() - t0) / 1e9, 2)
try:
    import subprocess
    q = subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max,pcie.link.width.max",
                        "--format=csv,noheader"], capture_output=True, text=True, timeout=20).stdout.strip()
    out["pcie_link_gen_width_current_max"] = q
except Exception as e:                                  # noqa: BLE001
    out["pcie_query_error"] = str(e)[:80]
print(json.dumps(out))
