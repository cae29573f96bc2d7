"""Deterministic synthetic clips.

Every bundled ``.mp4`` of the reference is a Git-LFS pointer, so tests and the
benchmark use textured synthetic frames: frame_0 = Gaussian-blurred uniform
noise (sigma 2.5) stretched to 0..255 per channel, frame_t = frame_0 warped by
the smooth field ``(u, v) = A*t*(1 + 0.3 sin(y/57), -0.5 + 0.3 cos(x/49))``
(A = 0.5 px/frame) plus N(0, 3^2) sensor noise.  Textured frames keep the
Farneback 2x2 systems well conditioned.

torch only (runs on CPU for tests and on the GPU for the benchmark).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _gauss_taps(sigma: float, device) -> torch.Tensor:
    r = int(math.ceil(3 * sigma))
    x = torch.arange(-r, r + 1, dtype=torch.float32, device=device)
    g = torch.exp(-(x * x) / (2 * sigma * sigma))
    return g / g.sum()


def synthetic_clip(n_frames: int, height: int, width: int, seed: int = 0,
                   device: str | torch.device = "cpu", amplitude: float = 0.5,
                   noise_sigma: float = 3.0) -> torch.Tensor:
    """Return uint8 BGR frames ``[n_frames, H, W, 3]`` on ``device``."""
    device = torch.device(device)
    gen = torch.Generator(device="cpu").manual_seed(seed)
    base = torch.rand((1, 3, height, width), generator=gen, dtype=torch.float32).to(device) * 255.0
    g = _gauss_taps(2.5, device)
    r = g.numel() // 2
    k = g.view(1, 1, 1, -1).repeat(3, 1, 1, 1)
    base = F.conv2d(F.pad(base, (r, r, 0, 0), mode="reflect"), k, groups=3)
    base = F.conv2d(F.pad(base, (0, 0, r, r), mode="reflect"), k.transpose(2, 3), groups=3)
    lo = base.amin(dim=(2, 3), keepdim=True)
    hi = base.amax(dim=(2, 3), keepdim=True)
    base = (base - lo) / (hi - lo) * 255.0

    ys = torch.arange(height, dtype=torch.float32, device=device).view(height, 1)
    xs = torch.arange(width, dtype=torch.float32, device=device).view(1, width)
    u = (1.0 + 0.3 * torch.sin(ys / 57.0)).expand(height, width)
    v = (-0.5 + 0.3 * torch.cos(xs / 49.0)).expand(height, width)

    out = torch.empty((n_frames, height, width, 3), dtype=torch.uint8, device=device)
    for t in range(n_frames):
        # sample frame_0 at (x - u*t, y - v*t): content moves by +(u, v) per frame
        sx = xs - amplitude * t * u
        sy = ys - amplitude * t * v
        gx = (sx + 0.5) / width * 2 - 1
        gy = (sy + 0.5) / height * 2 - 1
        grid = torch.stack([gx, gy], dim=-1).unsqueeze(0)
        fr = F.grid_sample(base, grid, mode="bilinear", padding_mode="reflection", align_corners=False)
        noise = torch.randn((1, 3, height, width), generator=gen, dtype=torch.float32).to(device)
        fr = (fr + noise_sigma * noise).clamp_(0, 255).round_()
        out[t] = fr[0].permute(1, 2, 0).to(torch.uint8)
    return out


def true_flow(height: int, width: int, amplitude: float = 0.5, device="cpu") -> torch.Tensor:
    """Ground-truth per-frame displacement ``[H, W, 2]`` of :func:`synthetic_clip`."""
    ys = torch.arange(height, dtype=torch.float32, device=device).view(height, 1)
    xs = torch.arange(width, dtype=torch.float32, device=device).view(1, width)
    u = amplitude * (1.0 + 0.3 * torch.sin(ys / 57.0)).expand(height, width)
    v = amplitude * (-0.5 + 0.3 * torch.cos(xs / 49.0)).expand(height, width)
    return torch.stack([u, v], dim=-1)
