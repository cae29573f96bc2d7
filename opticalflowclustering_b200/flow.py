"""Host side of the flow stage: Farneback plan, cv2-compatible call, visualisation.

Mirrors the cv2 calls the reference makes in
k-means-color-clustering/computeOpticalFlowModule.py:19-33 and
computeOpticalFlow.py:96-120.  torch is used for device memory and streams only;
all arithmetic is in libofc.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

_vp = C.c_void_p
OPTFLOW_USE_INITIAL_FLOW = 4          # cv2.OPTFLOW_USE_INITIAL_FLOW
OPTFLOW_FARNEBACK_GAUSSIAN = 256      # cv2.OPTFLOW_FARNEBACK_GAUSSIAN


def _stream_ptr() -> _vp:
    return _vp(torch.cuda.current_stream().cuda_stream)


def _ptr(t) -> _vp:
    return _vp(t.data_ptr()) if t is not None else _vp(0)


def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.OfcError("a CUDA device is required: this package has no CPU fallback")


def to_device_u8(a, device=None) -> torch.Tensor:
    """numpy / torch uint8 array -> contiguous CUDA tensor (no copy if already there)."""
    if isinstance(a, torch.Tensor):
        t = a
    else:
        a = np.asarray(a)
        if a.dtype != np.uint8:
            raise TypeError(f"expected uint8, got {a.dtype}")
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype != torch.uint8:
        raise TypeError(f"expected uint8, got {t.dtype}")
    if not t.is_cuda:
        _require_cuda()
        t = t.to(device or "cuda", non_blocking=True)
    return t.contiguous()


class FarnebackPlan:
    """Pyramid geometry, filter taps and workspace for one frame size.

    Parameters follow cv2.calcOpticalFlowFarneback; the defaults are the
    reference's literals (computeOpticalFlowModule.py:20-22).
    """

    def __init__(self, width: int, height: int, max_frames: int = 2, pyr_scale: float = 0.5, levels: int = 3,
                 winsize: int = 15, iterations: int = 3, poly_n: int = 5, poly_sigma: float = 1.2, flags: int = 0,
                 device=None):
        _require_cuda()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.width, self.height, self.max_frames = int(width), int(height), int(max_frames)
        self.params = (float(pyr_scale), int(levels), int(winsize), int(iterations), int(poly_n), float(poly_sigma), int(flags))
        L = _lib.lib()
        self._ptr = _vp()
        with torch.cuda.device(self.device):
            _lib.check(L.ofc_flow_plan_create(C.byref(self._ptr), self.width, self.height, self.max_frames,
                                              *self.params))
        self.workspace_bytes = int(L.ofc_flow_plan_workspace_bytes(self._ptr))
        self.workspace = torch.empty(self.workspace_bytes, dtype=torch.uint8, device=self.device)
        self.num_levels = int(L.ofc_flow_plan_num_levels(self._ptr))

    def keep_intermediates(self, keep: bool = True):
        """Also store the full-resolution pre-filtered image (``buffer(level, 0)``); fused away by default."""
        _lib.check(_lib.lib().ofc_flow_plan_keep_intermediates(self._ptr, int(bool(keep))))
        return self

    def level_size(self, level: int):
        w, h = C.c_int(), C.c_int()
        _lib.check(_lib.lib().ofc_flow_plan_level_size(self._ptr, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def buffer(self, level: int, kind: int, frame: int = 0) -> torch.Tensor:
        """View of an intermediate (kind 0 I, 1 RA, 2 RB, 3/4 flow ping/pong) for parity tests."""
        off, st = C.c_size_t(), C.c_size_t()
        _lib.check(_lib.lib().ofc_flow_plan_buffer(self._ptr, level, kind, C.byref(off), C.byref(st)))
        w, h = self.level_size(level)
        ch = {0: 1, 1: 4, 2: 1, 3: 2, 4: 2}[kind]
        b = self.workspace[off.value + frame * st.value: off.value + (frame + 1) * st.value]
        return b.view(torch.float32).view(h, w, ch)

    def sequence(self, gray: torch.Tensor, flow: torch.Tensor | None = None, minmax: torch.Tensor | None = None):
        """gray u8[n,H,W] (CUDA) -> flow f32[n-1,H,W,2]; optional minmax u32 bits [n-1,2]."""
        n = int(gray.shape[0])
        if tuple(gray.shape[1:]) != (self.height, self.width) or gray.dtype != torch.uint8 or not gray.is_cuda:
            raise ValueError(f"gray must be CUDA uint8 [n,{self.height},{self.width}], got {tuple(gray.shape)} {gray.dtype}")
        gray = gray.contiguous()
        if flow is None:
            flow = torch.empty((n - 1, self.height, self.width, 2), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ofc_farneback_sequence(self._ptr, _ptr(gray), n, _ptr(flow), _ptr(minmax),
                                                         _ptr(self.workspace), self.workspace_bytes, _stream_ptr()))
        return flow

    def pair(self, prev: torch.Tensor, nxt: torch.Tensor, flow: torch.Tensor | None = None,
             minmax: torch.Tensor | None = None, init_flow: torch.Tensor | None = None):
        """One pair, the literal cv2 call.  ``init_flow`` f32 [H,W,2] (CUDA) is required when the plan was
        created with cv2.OPTFLOW_USE_INITIAL_FLOW (4) and seeds the coarsest level."""
        for g in (prev, nxt):
            if tuple(g.shape) != (self.height, self.width) or g.dtype != torch.uint8 or not g.is_cuda:
                raise ValueError("prev/next must be CUDA uint8 [H,W] of the plan's size")
        prev, nxt = prev.contiguous(), nxt.contiguous()
        if flow is None:
            flow = torch.empty((self.height, self.width, 2), dtype=torch.float32, device=self.device)
        if self.params[6] & OPTFLOW_USE_INITIAL_FLOW:
            if (init_flow is None or tuple(init_flow.shape) != (self.height, self.width, 2) or init_flow.dtype != torch.float32
                    or not init_flow.is_cuda):
                raise ValueError("OPTFLOW_USE_INITIAL_FLOW needs init_flow: CUDA float32 [H,W,2] of the plan's size")
            init_flow = init_flow.contiguous()
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().ofc_farneback_pair_init(self._ptr, _ptr(prev), _ptr(nxt), _ptr(init_flow), _ptr(flow),
                                                              _ptr(minmax), _ptr(self.workspace), self.workspace_bytes,
                                                              _stream_ptr()))
            return flow
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ofc_farneback_pair(self._ptr, _ptr(prev), _ptr(nxt), _ptr(flow), _ptr(minmax),
                                                     _ptr(self.workspace), self.workspace_bytes, _stream_ptr()))
        return flow

    def stream_begin(self, first_gray: torch.Tensor):
        """Prime the streaming form with the first frame (CUDA uint8 [H,W]): its pre-filtered image and polynomial expansion
        stay in the workspace, so :meth:`stream_next` expands one frame per call (the reference's per-frame loop)."""
        g = first_gray.contiguous()
        if tuple(g.shape) != (self.height, self.width) or g.dtype != torch.uint8 or not g.is_cuda:
            raise ValueError("frame must be CUDA uint8 [H,W] of the plan's size")
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ofc_farneback_stream_begin(self._ptr, _ptr(g), _ptr(self.workspace), self.workspace_bytes,
                                                             _stream_ptr()))

    def stream_next(self, gray: torch.Tensor, flow: torch.Tensor | None = None, minmax: torch.Tensor | None = None):
        """Flow from the previous frame of the stream to ``gray`` (same bits as :meth:`pair`)."""
        g = gray.contiguous()
        if tuple(g.shape) != (self.height, self.width) or g.dtype != torch.uint8 or not g.is_cuda:
            raise ValueError("frame must be CUDA uint8 [H,W] of the plan's size")
        if flow is None:
            flow = torch.empty((self.height, self.width, 2), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ofc_farneback_stream_next(self._ptr, _ptr(g), _ptr(flow), _ptr(minmax), _ptr(self.workspace),
                                                            self.workspace_bytes, _stream_ptr()))
        return flow

    def __del__(self):
        try:
            if self._ptr:
                _lib.lib().ofc_flow_plan_destroy(self._ptr)
                self._ptr = _vp()
        except Exception:
            pass


_plan_cache: dict = {}


def _cached_plan(width, height, params, device) -> FarnebackPlan:
    key = (width, height, params, str(device))
    pl = _plan_cache.get(key)
    if pl is None:
        if len(_plan_cache) >= 4:
            _plan_cache.pop(next(iter(_plan_cache)))
        pl = FarnebackPlan(width, height, 2, *params, device=device)
        _plan_cache[key] = pl
    return pl


def calc_optical_flow_farneback(prev, next, flow=None, pyr_scale=0.5, levels=3, winsize=15, iterations=3,
                                poly_n=5, poly_sigma=1.2, flags=0):
    """Drop-in for ``cv2.calcOpticalFlowFarneback`` (same positional arguments).

    ``prev`` / ``next``: single-channel uint8 images of equal size, numpy or
    torch (CUDA).  Returns float32 ``[H, W, 2]`` of the same kind as the input.
    ``flags``: 0 (the reference's literal), ``cv2.OPTFLOW_FARNEBACK_GAUSSIAN`` (256) and/or
    ``cv2.OPTFLOW_USE_INITIAL_FLOW`` (4, ``flow`` is then the initial flow, float32 ``[H, W, 2]``);
    any other bit raises NotImplementedError (no CPU fallback).
    """
    as_numpy = not isinstance(prev, torch.Tensor)
    p = to_device_u8(prev)
    n = to_device_u8(next, p.device)
    if p.dim() != 2 or p.shape != n.shape:
        raise ValueError("prev and next must be single-channel images of equal size")
    params = (float(pyr_scale), int(levels), int(winsize), int(iterations), int(poly_n), float(poly_sigma), int(flags))
    plan = _cached_plan(int(p.shape[1]), int(p.shape[0]), params, p.device)
    init = None
    if int(flags) & OPTFLOW_USE_INITIAL_FLOW:
        if flow is None:
            raise ValueError("OPTFLOW_USE_INITIAL_FLOW needs the initial flow in `flow`")
        init = flow if isinstance(flow, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(flow, dtype=np.float32))
        init = init.to(device=p.device, dtype=torch.float32)
    out = plan.pair(p, n, init_flow=init)
    if as_numpy:
        res = out.cpu().numpy()
        if isinstance(flow, np.ndarray) and flow.shape == res.shape and flow.dtype == res.dtype:
            flow[...] = res
            return flow
        return res
    return out


def bgr2gray(bgr: torch.Tensor) -> torch.Tensor:
    """cv.cvtColor(..., COLOR_BGR2GRAY) on CUDA uint8 [..., 3] -> [...]"""
    bgr = bgr.contiguous()
    if bgr.shape[-1] != 3 or bgr.dtype != torch.uint8 or not bgr.is_cuda:
        raise ValueError("bgr must be a CUDA uint8 tensor [..., 3]")
    gray = torch.empty(bgr.shape[:-1], dtype=torch.uint8, device=bgr.device)
    with torch.cuda.device(bgr.device):
        _lib.check(_lib.lib().ofc_bgr2gray(_ptr(bgr), _ptr(gray), gray.numel(), _stream_ptr()))
    return gray


def flow_minmax(flow: torch.Tensor) -> torch.Tensor:
    """Per-frame IEEE bits of (min, max) of |flow|: int32 view of u32 [n, 2]."""
    f = flow.contiguous()
    n = int(f.shape[0])
    mm = torch.empty((n, 2), dtype=torch.int32, device=f.device)
    with torch.cuda.device(f.device):
        _lib.check(_lib.lib().ofc_flow_minmax(_ptr(f), n, f.shape[1] * f.shape[2], _ptr(mm), _stream_ptr()))
    return mm


def flow_to_bgr(flow: torch.Tensor, minmax: torch.Tensor | None = None, want_mean_magnitude: bool = False,
                want_hsv: bool = False):
    """cartToPolar + hue byte + NORM_MINMAX + HSV2BGR (computeOpticalFlowModule.py:25-33).

    flow f32 [n,H,W,2] (CUDA) -> BGR u8 [n,H,W,3] (and mean |flow| per frame f64[n]).
    """
    f = flow.contiguous()
    if f.dim() != 4 or f.shape[-1] != 2 or f.dtype != torch.float32 or not f.is_cuda:
        raise ValueError("flow must be a CUDA float32 tensor [n,H,W,2]")
    n, H, W = int(f.shape[0]), int(f.shape[1]), int(f.shape[2])
    if minmax is None:
        minmax = flow_minmax(f)
    bgr = torch.empty((n, H, W, 3), dtype=torch.uint8, device=f.device)
    if want_hsv:
        # also the HSV `mask` of the reference (H, 255, V): returns (bgr, hsv)
        hsv = torch.empty_like(bgr)
        with torch.cuda.device(f.device):
            _lib.check(_lib.lib().ofc_flow_to_hsv(_ptr(f), n, H, W, _ptr(minmax), _ptr(bgr), _ptr(hsv), _stream_ptr()))
        return bgr, hsv
    mag = torch.empty(n, dtype=torch.float64, device=f.device) if want_mean_magnitude else None
    with torch.cuda.device(f.device):
        _lib.check(_lib.lib().ofc_flow_to_bgr(_ptr(f), n, H, W, _ptr(minmax), _ptr(bgr), _ptr(mag), _stream_ptr()))
    if want_mean_magnitude:
        return bgr, mag / float(H * W)
    return bgr
