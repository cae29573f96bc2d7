"""CPU oracle: 8-bit colour / visualisation primitives of the flow path.

TEST INFRASTRUCTURE ONLY (see oracle/farneback_np.py header for the rule).

Restates, formula by formula, what the reference's
``ComputeOpticalFLow.compute`` (k-means-color-clustering/
computeOpticalFlowModule.py:18-36) obtains from cv2 4.13.0:
  :19  cv.cvtColor(BGR2GRAY)      -> bgr2gray
  :25  cv.cartToPolar             -> cart_to_polar
  :28  mask[...,0] = angle*180/pi/2 (float32 -> uint8 truncation) -> hue_byte
  :31  cv.normalize(NORM_MINMAX)  -> normalize_minmax_u8
  :33  cv.cvtColor(HSV2BGR)       -> hsv2bgr_s255
and, for the grid stage (KmeanGrids.py:86-92,336, color_kmeans.py:121),
  cv2.cvtColor(BGR2HSV) on 8-bit -> bgr2hsv_u8.

Pinned by: tests/test_oracle_viz.py against live cv2 (exhaustive for the
integer formulas) and tests/golden/viz_*.npz.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def bgr2gray(bgr: np.ndarray) -> np.ndarray:
    b = bgr[..., 0].astype(np.int64)
    g = bgr[..., 1].astype(np.int64)
    r = bgr[..., 2].astype(np.int64)
    return ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(np.uint8)


def _fma32(a, b, c):
    """float32 fused multiply-add emulated in float64.

    a*b is exact in float64 for float32 inputs (48-bit product); the following
    add and the final rounding to float32 are a double rounding, which differs
    from a true fmaf only in rare half-way cases.
    """
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F32)


_ATAN_P = [F32(np.float32(v) * np.float32(180.0 / np.pi)) for v in
           (0.9997878412794807, -0.3258083974640975, 0.1555786518463281, -0.04432655554792128)]


def cart_to_polar(x: np.ndarray, y: np.ndarray):
    """cv::cartToPolar(x, y) -> magnitude, angle in radians [0, 2pi)."""
    x = x.astype(F32)
    y = y.astype(F32)
    mag = np.sqrt(_fma32(x, x, (y * y).astype(F32))).astype(F32)
    ax = np.abs(x)
    ay = np.abs(y)
    mx = np.maximum(ax, ay)
    mn = np.minimum(ax, ay)
    c = (mn / (mx + F32(np.finfo(np.float64).eps))).astype(F32)
    c2 = (c * c).astype(F32)
    p1, p3, p5, p7 = _ATAN_P
    # cv2 4.13 evaluates the polynomial with fused multiply-adds (v_fma): this
    # form is bit-exact against cv2.cartToPolar on 2 M random samples, the
    # separate mul/add form only 99.5 %.
    k = lambda v: np.full_like(c2, v)
    a = _fma32(_fma32(_fma32(c2, k(p7), k(p5)), c2, k(p3)), c2, k(p1))
    a = (a * c).astype(F32)
    a = np.where(ax < ay, F32(90) - a, a).astype(F32)
    a = np.where(x < 0, F32(180) - a, a).astype(F32)
    a = np.where(y < 0, F32(360) - a, a).astype(F32)
    ang = (a * F32(np.pi / 180.0)).astype(F32)
    return mag, ang


def hue_byte(angle_rad: np.ndarray) -> np.ndarray:
    """mask[...,0] = angle*180/np.pi/2 : float32 expression, C truncation to u8."""
    v = angle_rad.astype(F32) * F32(180)
    v = (v / F32(np.pi)).astype(F32)
    v = (v / F32(2)).astype(F32)
    return v.astype(np.int32).astype(np.uint8)


def normalize_minmax_u8(mag: np.ndarray) -> np.ndarray:
    """cv.normalize(mag, None, 0, 255, NORM_MINMAX) stored into a uint8 array.

    cv2 4.13 evaluates ``fmaf(src, scale_f, shift_f)`` with ``scale_f =
    f32(255/(max-min))`` and ``shift_f = -(min * scale_f)`` rounded in float32
    (the shift is derived from the *rounded* scale): 3000/3000 random arrays
    bit-exact; deriving the shift from the double scale matches only ~74 %.
    """
    mag = mag.astype(F32)
    mn32 = mag.min()
    mx32 = mag.max()
    rng = float(mx32) - float(mn32)
    scale = 255.0 * (1.0 / rng if rng > np.finfo(np.float64).eps else 0.0)
    scale_f = F32(scale)
    shift_f = F32(-(mn32 * scale_f))
    out = _fma32(mag, np.full_like(mag, scale_f), np.full_like(mag, shift_f))
    return out.astype(np.int32).astype(np.uint8)


_SECTOR = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])


def hsv2bgr_s255(h: np.ndarray, v: np.ndarray) -> np.ndarray:
    """cv.cvtColor(HSV2BGR) on 8-bit input (H range 180) with S = 255 (the only
    saturation this path produces, computeOpticalFlowModule.py:15).

    float32 sector formula.  cv2 4.13 converts each image row in 32-pixel SIMD
    groups whose float->u8 step TRUNCATES, and finishes the last ``W % 32``
    pixels of the row with scalar code that ROUNDS (half to even); verified
    exhaustively over H in 0..180, V in 0..255 for both paths and over row
    widths 31..1000.  The last axis of ``h`` / ``v`` is taken as the image row.
    """
    hf = h.astype(F32) * F32(6.0 / 180.0)
    vf = v.astype(F32) * F32(1.0 / 255.0)
    sf = np.full(h.shape, 255, np.uint8).astype(F32) * F32(1.0 / 255.0)
    sec = np.floor(hf).astype(np.int32)
    fr = (hf - sec.astype(F32)).astype(F32)
    sec = np.mod(sec, 6)
    one = F32(1)
    tab = np.stack([
        vf,
        (vf * (one - sf)).astype(F32),
        (vf * (one - (sf * fr).astype(F32))).astype(F32),
        (vf * (one - (sf * (one - fr)).astype(F32))).astype(F32),
    ], axis=-1)
    idx = _SECTOR[sec]                                   # [...,3] -> b,g,r tab index
    bgr = np.take_along_axis(tab, idx, axis=-1)
    out = (bgr * F32(255)).astype(F32)
    W = h.shape[-1] if h.ndim else 1
    tail = np.arange(W) >= W - (W % 32)                  # scalar tail of each row: rounded
    res = np.where(tail.reshape((1,) * (h.ndim - 1) + (W, 1)) if h.ndim else tail, np.rint(out), np.trunc(out))
    return np.clip(res, 0, 255).astype(np.uint8)


_HDIV = np.zeros(256, np.int64)
_SDIV = np.zeros(256, np.int64)
for _i in range(1, 256):
    _HDIV[_i] = int(np.rint((180 << 12) / (6.0 * _i)))
    _SDIV[_i] = int(np.rint((255 << 12) / (1.0 * _i)))


def bgr2hsv_u8(bgr: np.ndarray) -> np.ndarray:
    """cv.cvtColor(BGR2HSV) on 8-bit input, H in 0..179 (integer table formula)."""
    b = bgr[..., 0].astype(np.int64)
    g = bgr[..., 1].astype(np.int64)
    r = bgr[..., 2].astype(np.int64)
    v = np.maximum(np.maximum(b, g), r)
    mn = np.minimum(np.minimum(b, g), r)
    d = v - mn
    h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * d, r - g + 4 * d))
    h = (h * _HDIV[d] + (1 << 11)) >> 12
    h = np.where(h < 0, h + 180, h)
    s = (d * _SDIV[v] + (1 << 11)) >> 12
    return np.stack([h, s, v], axis=-1).astype(np.uint8)


def flow_to_bgr(flow: np.ndarray):
    """computeOpticalFlowModule.py:25-33 given a flow field -> (BGR u8, magnitude)."""
    mag, ang = cart_to_polar(flow[..., 0], flow[..., 1])
    hb = hue_byte(ang)
    vb = normalize_minmax_u8(mag)
    return hsv2bgr_s255(hb, vb), mag
