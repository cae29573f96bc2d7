"""Experiment: do two ClipPipelines on two streams (alternate chunks) finish a clip sooner than one?  The step is a
chain of kernels of very different character (latency-bound walks that own every register of the SM, issue-bound
stage-1 kernels, small coarse levels); with two independent chains in flight the hardware can fill one chain's gaps and
tails with the other's CTAs.  Prints pairs/s for 1 and 2 lanes (CUDA events, 24 steps after 6 warm-up steps)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opticalflowclustering_b200.pipeline import ClipPipeline
from opticalflowclustering_b200.synthetic import synthetic_clip

H, W, F = 1080, 1920, int(os.environ.get("OFC_CHUNK", "33"))
P = F - 1
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
T = 129
clip = synthetic_clip(T, H, W, seed=0, device=dev)
starts = [(i * P) % (T - F + 1) for i in range(64)]


def run(lanes, steps=24, warm=6):
    pipes = [ClipPipeline(W, H, chunk_frames=F, device=dev) for _ in range(lanes)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(lanes)]
    main = torch.cuda.current_stream()

    def body(i):
        l = i % lanes
        with torch.cuda.stream(streams[l]):
            pipes[l].run_chunk(clip[starts[i]:starts[i] + F])

    for i in range(warm):
        body(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for s in streams:
        s.wait_event(e0)
    for i in range(warm, warm + steps):
        body(i)
    for s in streams:
        ev = torch.cuda.Event()
        ev.record(s)
        main.wait_event(ev)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{lanes} lane(s): {ms:.3f} ms/step  {P / ms * 1e3:.1f} pairs/s", flush=True)
    del pipes
    torch.cuda.empty_cache()


run(1)
run(2)
run(3)
run(1)
