"""Short, fixed launch sequence for ncu: 2 warm steps + 1 profiled step of the 1080p pipeline (OFC_CHUNK frames, default 9)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opticalflowclustering_b200.pipeline import ClipPipeline
from opticalflowclustering_b200.synthetic import synthetic_clip

H, W = {"720p": (720, 1280), "1080p": (1080, 1920), "4k": (2160, 3840)}[os.environ.get("OFC_SIZE", "1080p")]
F = int(os.environ.get("OFC_CHUNK", "9"))
clip = synthetic_clip(F, H, W, seed=0, device="cuda")
pipe = ClipPipeline(W, H, chunk_frames=F, n_clusters=int(os.environ.get("OFC_K", "1")))
for _ in range(2):
    pipe.run_chunk(clip)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()          # `ncu --profile-from-start off` captures exactly the third step
pipe.run_chunk(clip)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok", pipe.km_hue[0, :8].tolist())
