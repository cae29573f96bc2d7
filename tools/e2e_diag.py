"""Where does the end-to-end loop lose time against the device-resident loop?  Times the 1080p step in variants that add
one ingredient of bench.py's e2e loop at a time (CUDA events over 20 steps each, after 5 warm-up steps)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opticalflowclustering_b200.pipeline import ClipPipeline
from opticalflowclustering_b200.synthetic import synthetic_clip

H, W, F = 1080, 1920, 33
P = F - 1
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
T = 129
clip = synthetic_clip(T, H, W, seed=0, device=dev)
pipe = ClipPipeline(W, H, chunk_frames=F, device=dev)
starts = [(i * P) % (T - F + 1) for i in range(64)]
host = clip.cpu().pin_memory()
stage = [torch.empty((P, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(3)]
res = [torch.empty((P, 350), dtype=torch.uint8).pin_memory() for _ in range(2)] + [torch.empty(P, dtype=torch.float64).pin_memory()]


def timed(name, body, steps=20, warm=5):
    for i in range(warm):
        body(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(warm, warm + steps):
        body(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{name:44s} {ms:7.3f} ms/step  {P / ms * 1e3:8.1f} pairs/s", flush=True)


timed("resident chunk of 33 frames", lambda i: pipe.run_chunk(clip[starts[i]:starts[i] + F]))
pipe.run_chunk(clip[0:2])
timed("carry, 32 new resident frames", lambda i: pipe.run_chunk(clip[starts[i] + 1:starts[i] + F], carry=True))


def carry_nocopy(i):
    pipe._last_gray_index = 0                  # timing only: skips the device copy of the carried gray frame (wrong pairs 0)
    pipe.run_chunk(clip[starts[i] + 1:starts[i] + F], carry=True)


timed("carry without the gray-frame copy (timing only)", carry_nocopy)
timed("resident chunk of 33 frames (again)", lambda i: pipe.run_chunk(clip[starts[i]:starts[i] + F]))
pipe.run_chunk(clip[0:2])


def with_d2h(i):
    pipe.run_chunk(clip[starts[i] + 1:starts[i] + F], carry=True)
    res[0].copy_(pipe.avg_hue[:P], non_blocking=True)
    res[1].copy_(pipe.km_hue[:P], non_blocking=True)
    res[2].copy_(pipe.mag_sum[:P], non_blocking=True)


timed("carry + 3 in-line result read-backs", with_d2h)

copy_stream = torch.cuda.Stream(device=dev)
ready = [torch.cuda.Event() for _ in range(3)]
freed = [torch.cuda.Event() for _ in range(3)]
main = torch.cuda.current_stream()
for b in range(3):
    freed[b].record(main)


def staged(i, upload=True):
    b = i % 3
    with torch.cuda.stream(copy_stream):
        copy_stream.wait_event(freed[b])
        if upload:
            c = starts[i] + 1
            stage[b].copy_(host[c:c + P], non_blocking=True)
        ready[b].record(copy_stream)
    main.wait_event(ready[b])
    pipe.run_chunk(stage[b], carry=True)
    freed[b].record(main)


for b in range(3):
    stage[b].copy_(clip[1:F])
timed("carry from staging buffers, events, no upload", lambda i: staged(i, False))
timed("carry from staging buffers, upload same step", lambda i: staged(i, True))
