"""Drop-in for the reference's k-means-color-clustering/computeOpticalFlowModule.py.

Same class name (typo included), constructor, attributes and ``compute``
contract (computeOpticalFlowModule.py:6-36): stateful per video, keeps
``prev_gray``; ``compute(frame)`` returns a *new* uint8 BGR visualisation.
numpy in -> numpy out; CUDA torch tensors in -> CUDA torch tensors out.
"""
from __future__ import annotations

import numpy as np
import torch

from . import flow as _flow


class ComputeOpticalFLow:
    def __init__(self, firstframe):
        self.firstframe = firstframe
        self.width = int(firstframe.shape[1])
        self.height = int(firstframe.shape[0])
        self._numpy = not isinstance(firstframe, torch.Tensor)
        if self._numpy:
            self.outputImg = np.zeros([self.height, 2 * self.width, 3], dtype=firstframe.dtype)
            self._mask0 = np.zeros_like(firstframe)
            self._mask0[..., 1] = 255
        else:
            self.outputImg = None
            self._mask0 = torch.zeros_like(firstframe)
            self._mask0[..., 1] = 255
        self._mask = None                # HSV image of the latest pair, produced on demand (see `mask`)
        dev_frame = _flow.to_device_u8(firstframe)
        self._plan = _flow.FarnebackPlan(self.width, self.height, 2, 0.5, 3, 15, 3, 5, 1.2, 0, device=dev_frame.device)
        self._prev_gray = _flow.bgr2gray(dev_frame)
        # the previous frame's pre-filtered image and polynomial expansion stay on the device: compute() expands one frame
        self._plan.stream_begin(self._prev_gray)
        self._minmax = torch.empty((1, 2), dtype=torch.int32, device=dev_frame.device)
        self.last_flow = None            # CUDA float32 [H,W,2] of the latest pair

    @property
    def prev_gray(self):
        return self._prev_gray.cpu().numpy() if self._numpy else self._prev_gray

    @property
    def mask(self):
        """The reference's ``self.mask`` (computeOpticalFlowModule.py:14-15, 28-31): HSV image with H = direction byte,
        S = 255, V = min-max-normalised magnitude byte of the latest pair (all zero with S = 255 before the first
        ``compute``).  ``compute`` itself only needs the BGR image, so the HSV bytes are produced on first access."""
        if self.last_flow is None:
            return self._mask0
        if self._mask is None:
            _, hsv = _flow.flow_to_bgr(self.last_flow.unsqueeze(0), self._minmax, want_hsv=True)
            self._mask = hsv[0].cpu().numpy() if self._numpy else hsv[0]
        return self._mask

    def compute(self, frame):
        dev_frame = _flow.to_device_u8(frame, self._prev_gray.device)
        gray = _flow.bgr2gray(dev_frame)
        fl = self._plan.stream_next(gray, minmax=self._minmax)
        bgr = _flow.flow_to_bgr(fl.unsqueeze(0), self._minmax)[0]
        self._prev_gray = gray
        self.last_flow = fl
        self._mask = None
        if self._numpy:
            return bgr.cpu().numpy()
        return bgr
