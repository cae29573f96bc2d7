"""Host side of the cosine stage.

Mirrors the reference's two scripts
  k-means-color-clustering/findCosineDifferentVectors.py:5-66   (sliding window cosine)
  k-means-color-clustering/computeVectorDistance.py:22-43      (cosine, "row 0" quirk, L1)
plus the row-vs-query form used for the 1M-vector sweep of BASELINE.json.  All arithmetic
is in libofc.so (float64; the reference's hue vectors are integers, so dot products and
squared norms are exact and the quotient has numpy's bits).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import kmeans as _km
from .kmeans import _Ctx, _DT, _ptr, _target_device


def _vec(a, device):
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a)))
    return t.reshape(-1).to(device=device, dtype=torch.float64).contiguous()


def _device():
    return _target_device(None)


def sliding_cosine(short, long_, return_sims: bool = False):
    """Loop of findCosineDifferentVectors.py:48-61: cosine of ``short`` against every window
    of ``long_``; returns ``(max similarity, LAST arg-max index)`` (and the similarities).
    ``len(long_) < len(short)`` gives the loop's initial state ``(-1, -1)``."""
    ctx = _Ctx(_device() if not (isinstance(short, torch.Tensor) and short.is_cuda) else short.device)
    a, b = _vec(short, ctx.device), _vec(long_, ctx.device)
    n, m = a.numel(), b.numel()
    if n == 0 or m < n:
        return (-1, -1, np.empty(0)) if return_sims else (-1, -1)
    sims = torch.empty(m - n + 1, dtype=torch.float64, device=ctx.device)
    best = torch.empty(1, dtype=torch.float64, device=ctx.device)
    idx = torch.empty(1, dtype=torch.int64, device=ctx.device)
    ctx.check(ctx.lib.ofc_sliding_cosine(_ptr(a), int(n), _ptr(b), C.c_int64(m), _ptr(sims), _ptr(best), _ptr(idx),
                                         ctx.stream()))
    res = (float(best.item()), int(idx.item()))
    return res + (sims.cpu().numpy(),) if return_sims else res


def calculate_cosine_similarity(file1_hue, nobounce_hue):
    """Same name and contract as findCosineDifferentVectors.py:5-26 (0 when a norm is 0)."""
    a = np.asarray(file1_hue).reshape(-1)
    b = np.asarray(nobounce_hue).reshape(-1)
    if a.size != b.size:
        raise ValueError(f"shapes {a.shape} and {b.shape} not aligned")       # np.dot's error
    best, _ = sliding_cosine(a, b)
    if best == 0.0 and (not a.any() or not b.any()):
        return 0                                                            # the reference returns int 0
    return np.float64(best)


def row_cosine(X, q):
    """``out[i] = cos(X[i, :], q)`` for ``X [N, D]`` (uint8 / float32 / float64) -> float64 ``[N]``."""
    if isinstance(X, np.ndarray):
        X = torch.from_numpy(np.ascontiguousarray(X))
    if not X.is_cuda and _km._TEST_LIBRARY is None:
        X = X.to(_target_device(X))
    if X.dtype not in _DT:
        X = X.to(torch.float64)
    X = X.contiguous()
    ctx = _Ctx(X.device)
    qv = _vec(q, ctx.device)
    if X.dim() != 2 or qv.numel() != X.shape[1]:
        raise ValueError("X must be [N, D] and q [D]")
    out = torch.empty(X.shape[0], dtype=torch.float64, device=ctx.device)
    ctx.check(ctx.lib.ofc_row_cosine(_ptr(X), _DT[X.dtype], C.c_int64(X.shape[0]), int(X.shape[1]), _ptr(qv), _ptr(out),
                                     ctx.stream()))
    return out


def vector_distance(hsv1, hsv2):
    """computeVectorDistance.py:22-43 on two hue columns -> ``(similarity [[c]], row, distance)``:
    sklearn's cosine of the flattened vectors, the row the script prints as "Cosine similarity"
    (``hsv1[0]*hsv2[j] / (|hsv1[j]|*|hsv2[j]|)``), and ``sum_i |hsv1[i]-hsv2[i]|`` over the
    common prefix."""
    ctx = _Ctx(_device())
    a, b = _vec(hsv1, ctx.device), _vec(hsv2, ctx.device)
    if a.numel() != b.numel():
        # np.dot(hsv1, hsv2.T) works for unequal lengths but the division by the two norm vectors
        # does not broadcast: the script dies there (computeVectorDistance.py:25)
        raise ValueError(f"operands could not be broadcast together with shapes ({a.numel()},) ({b.numel()},) ")
    n = a.numel()
    cos = torch.empty(1, dtype=torch.float64, device=ctx.device)
    row = torch.empty(n, dtype=torch.float64, device=ctx.device)
    l1 = torch.empty(1, dtype=torch.float64, device=ctx.device)
    ctx.check(ctx.lib.ofc_vector_distance(_ptr(a), _ptr(b), C.c_int64(n), _ptr(cos), _ptr(row), _ptr(l1), ctx.stream()))
    return np.array([[float(cos.item())]]), row.cpu().numpy(), float(l1.item())
