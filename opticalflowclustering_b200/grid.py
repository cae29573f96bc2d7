"""Host side of the grid stage (per-cell means, hues and the k = 1 colour cluster).

Mirrors overlayGridAndComputeAvgColor of the reference
(k-means-color-clustering/KmeanGrids.py:52-113, drawGridsAndOutputCSV.py:47-135)
and the per-cell preprocess_image + KMeans(n_clusters=1) of
KmeanGrids.py:269-339.  All arithmetic is integer and runs in libofc.so.
"""
from __future__ import annotations

import torch

from . import _lib
from .flow import _ptr, _stream_ptr

DEFAULT_GRID = {'rows': 14, 'cols': 25, 'cell_width': 50, 'cell_height': 50}   # KmeanGrids.py:177


def grid_cells(bgr: torch.Tensor, rows: int = 14, cols: int = 25, draw_lines: bool = True, threshold: int = 30,
               want=("avg_bgr", "avg_hue", "km_centre", "km_hue")) -> dict:
    """One CTA per (cell, frame) over CUDA uint8 frames ``[n, H, W, 3]``.

    Returns a dict of CUDA uint8 tensors (cells = rows*cols, reference order):
      avg_bgr [n,cells,3]  floor(mean) per channel   (np.mean(...).astype(uint8))
      avg_hue [n,cells]    cv2 BGR2HSV hue of avg_bgr
      km_centre [n,cells,4] np.rint of the KMeans(1) centre of (c0,c1,c2,alpha)
      km_hue  [n,cells]    hue of km_centre[:3]
      km_sums [n,cells,4]  uint32 channel sums of the k-means input (optional)
    """
    if bgr.dim() == 3:
        bgr = bgr.unsqueeze(0)
    bgr = bgr.contiguous()
    if bgr.dim() != 4 or bgr.shape[-1] != 3 or bgr.dtype != torch.uint8 or not bgr.is_cuda:
        raise ValueError("bgr must be a CUDA uint8 tensor [n,H,W,3]")
    n, H, W = int(bgr.shape[0]), int(bgr.shape[1]), int(bgr.shape[2])
    cells = rows * cols
    dev = bgr.device
    shapes = {"avg_bgr": ((n, cells, 3), torch.uint8), "avg_hue": ((n, cells), torch.uint8),
              "km_centre": ((n, cells, 4), torch.uint8), "km_hue": ((n, cells), torch.uint8),
              "km_sums": ((n, cells, 4), torch.int32)}
    out = {k: torch.empty(shapes[k][0], dtype=shapes[k][1], device=dev) for k in want}
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().ofc_grid_cells(_ptr(bgr), n, H, W, rows, cols, int(bool(draw_lines)), int(threshold),
                                             _ptr(out.get("avg_bgr")), _ptr(out.get("avg_hue")),
                                             _ptr(out.get("km_centre")), _ptr(out.get("km_hue")),
                                             _ptr(out.get("km_sums")), _stream_ptr()))
    return out


def grid_kmeans_cells(bgr: torch.Tensor, n_clusters: int, rows: int = 14, cols: int = 25, draw_lines: bool = True,
                      threshold: int = 30, seed: int = 0, first_frame: int = 0, max_iter: int = 300, tol: float = 1e-4,
                      swap_rb: bool = False, want_centres: bool = False) -> dict:
    """The reference's per-cell ``preprocess_image`` + ``KMeans(n_clusters).fit`` + dominant cluster -> ``np.rint`` ->
    hue (KmeanGrids.py:269-339, 376-392) for every cell of CUDA uint8 frames ``[n, H, W, 3]`` in one launch: the cells
    are gathered inside the k-means kernel, nothing but the results leaves the SMs.  k-means++ seeding from ``seed``
    (problem index ``(first_frame + frame) * cells + cell``).  Returns CUDA tensors ``dom_centre`` u8 [n,cells,4],
    ``dom_hue`` u8 [n,cells], ``n_iter`` i32 [n,cells] (+ ``centres`` f64 [n,cells,k,4], ``counts`` i64 [n,cells,k])."""
    import ctypes as C
    if bgr.dim() == 3:
        bgr = bgr.unsqueeze(0)
    bgr = bgr.contiguous()
    if bgr.dim() != 4 or bgr.shape[-1] != 3 or bgr.dtype != torch.uint8 or not bgr.is_cuda:
        raise ValueError("bgr must be a CUDA uint8 tensor [n,H,W,3]")
    n, H, W = int(bgr.shape[0]), int(bgr.shape[1]), int(bgr.shape[2])
    cells, k, dev = rows * cols, int(n_clusters), bgr.device
    L = _lib.lib()
    out = {"dom_centre": torch.empty((n, cells, 4), dtype=torch.uint8, device=dev),
           "dom_hue": torch.empty((n, cells), dtype=torch.uint8, device=dev),
           "n_iter": torch.empty((n, cells), dtype=torch.int32, device=dev)}
    if want_centres:
        out["centres"] = torch.empty((n, cells, k, 4), dtype=torch.float64, device=dev)
        out["counts"] = torch.empty((n, cells, k), dtype=torch.int64, device=dev)
    ws = torch.empty(max(8, int(L.ofc_grid_kmeans_cells_workspace_bytes(n, H, W, rows, cols, k))), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.ofc_grid_kmeans_cells(_ptr(bgr), n, H, W, rows, cols, int(bool(draw_lines)), int(threshold), int(bool(swap_rb)),
                                           k, C.c_uint64(int(seed) & (2 ** 64 - 1)), C.c_uint64(int(first_frame)), int(max_iter),
                                           C.c_double(float(tol)), _ptr(out["dom_centre"]), _ptr(out["dom_hue"]),
                                           _ptr(out.get("centres")), _ptr(out.get("counts")), _ptr(out["n_iter"]), _ptr(ws),
                                           C.c_size_t(ws.numel()), _stream_ptr()))
    return out


def draw_grid(bgr: torch.Tensor, rows: int = 14, cols: int = 25) -> torch.Tensor:
    """In place: the white 1-px rectangles of KmeanGrids.py:108 on CUDA uint8 [n,H,W,3]."""
    if bgr.dim() == 3:
        view = bgr.unsqueeze(0)
    else:
        view = bgr
    if not view.is_contiguous() or view.dtype != torch.uint8 or not view.is_cuda or view.shape[-1] != 3:
        raise ValueError("bgr must be a contiguous CUDA uint8 tensor [n,H,W,3]")
    n, H, W = int(view.shape[0]), int(view.shape[1]), int(view.shape[2])
    with torch.cuda.device(view.device):
        _lib.check(_lib.lib().ofc_draw_grid(_ptr(view), n, H, W, rows, cols, _stream_ptr()))
    return bgr
