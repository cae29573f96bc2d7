"""CPU oracle: grid-cell aggregation, per-cell k-means, cosine scripts.

TEST INFRASTRUCTURE ONLY (see oracle/farneback_np.py header for the rule).

Restates the reference's own Python glue (no third-party arithmetic beyond the
8-bit colour formulas in oracle/viz_np.py and the KMeans restatement in
oracle/kmeans_np.py):
  grid_cells / grid_mean_hues   KmeanGrids.py:52-113, drawGridsAndOutputCSV.py:47-135
  draw_grid                     cv2.rectangle(frame,(x1,y1),(x2,y2),(255,255,255),1), :108
  preprocess_image              KmeanGrids.py:269-286, color_kmeans.py:35-52
  cluster_colors_k1 / _general  KmeanGrids.py:288-339, color_kmeans.py:54-135
  sliding_cosine                findCosineDifferentVectors.py:5-66
  vector_distance               computeVectorDistance.py:22-43

Pinned by the reference's goldens G1-G6, copied in part to
tests/golden/ by tests/golden/make_golden.py.
"""
from __future__ import annotations

import numpy as np

from . import kmeans_np, viz_np


def grid_cells(height: int, width: int, rows: int, cols: int):
    """[(x1, y1, x2, y2)] in the reference's row-major cell order."""
    x_step = int(width / cols)
    y_step = int(height / rows)
    cells = []
    for y in range(rows):
        for x in range(cols):
            x1 = x * x_step
            y1 = y * y_step
            cells.append((x1, y1, min(x1 + x_step, width), min(y1 + y_step, height)))
    return cells


def draw_rect_1px(frame: np.ndarray, x1, y1, x2, y2, value=255):
    """cv2.rectangle thickness 1 (inclusive corners, clipped to the image)."""
    h, w = frame.shape[:2]
    xs = slice(max(x1, 0), min(x2, w - 1) + 1)
    ys = slice(max(y1, 0), min(y2, h - 1) + 1)
    if 0 <= y1 < h:
        frame[y1, xs] = value
    if 0 <= y2 < h:
        frame[y2, xs] = value
    if 0 <= x1 < w:
        frame[ys, x1] = value
    if 0 <= x2 < w:
        frame[ys, x2] = value


def grid_mean_hues(frame: np.ndarray, rows: int = 14, cols: int = 25):
    """overlayGridAndComputeAvgColor: per-cell floor(mean BGR) and its hue.

    Mutates ``frame`` exactly like the reference (white 1-px rectangles drawn
    after each cell's mean).  Returns (avg_bgr u8[n,3], hue u8[n],
    rois list of views into ``frame``).
    """
    h, w = frame.shape[:2]
    avgs, hues, rois = [], [], []
    for (x1, y1, x2, y2) in grid_cells(h, w, rows, cols):
        roi = frame[y1:y2, x1:x2]
        n = roi.shape[0] * roi.shape[1]
        s = roi.reshape(n, -1).astype(np.int64).sum(axis=0)
        avg = (s // n).astype(np.uint8)                  # np.mean(...).astype(uint8)
        avgs.append(avg)
        hues.append(viz_np.bgr2hsv_u8(avg[None, None, :])[0, 0, 0])
        draw_rect_1px(frame, x1, y1, x2, y2)
        rois.append(roi)
    return np.array(avgs, np.uint8), np.array(hues, np.uint8), rois


def preprocess_image(image: np.ndarray) -> np.ndarray:
    """In-place ``image[image<30]=0``; append alpha = 255*(gray>0)."""
    image[image < 30] = 0
    gray = viz_np.bgr2gray(image)
    alpha = np.where(gray > 0, 255, 0).astype(np.uint8)
    return np.concatenate([image, alpha[..., None]], axis=-1)


def cluster_colors_k1(image4: np.ndarray):
    """cluster_colors with n_clusters=1 -> (np.rint(centre)[4] float64, hue)."""
    flat = image4.reshape(-1, 4)
    n = flat.shape[0]
    s = flat.astype(np.int64).sum(axis=0)
    c = kmeans_np.rint_mean_exact(s, n).astype(np.float64)
    rgb0 = np.array([[[c[0], c[1], c[2]]]], dtype=np.uint8)
    hue = viz_np.bgr2hsv_u8(rgb0)[0, 0, 0]
    return c, hue


def cluster_colors_general(image4: np.ndarray, init: np.ndarray):
    """cluster_colors with k>1 and an explicit initial centre set.

    The reference leaves random_state unset (KmeanGrids.py:300) so k>1 is only
    defined up to the initial centres; with ``init`` given this follows
    fit -> predict -> bincount -> stable sort by share (descending) ->
    np.rint(top centre) -> BGR2HSV hue.
    """
    flat = image4.reshape(-1, 4)
    labels, centers, inertia, n_iter = kmeans_np.kmeans_fit(flat, init)
    pred = kmeans_np.kmeans_predict(flat, centers)
    counts = np.bincount(pred, minlength=centers.shape[0])
    share = counts.astype(float) / len(flat)
    order = sorted(range(len(share)), key=lambda i: share[i], reverse=True)   # stable
    c = np.rint(centers[order[0]])
    rgb0 = np.array([[[c[0], c[1], c[2]]]], dtype=np.uint8)
    hue = viz_np.bgr2hsv_u8(rgb0)[0, 0, 0]
    return c, hue, labels, centers, inertia, n_iter


def cosine_similarity(a, b):
    """calculate_cosine_similarity (findCosineDifferentVectors.py:5-26)."""
    a = np.asarray(a)
    b = np.asarray(b)
    na = np.linalg.norm(a)
    nb = np.linalg.norm(b)
    if na == 0 or nb == 0:
        return 0
    return np.dot(a, b) / (na * nb)


def sliding_cosine(short, long_):
    """Loop of findCosineDifferentVectors.py:48-61 -> (max similarity, last arg-max)."""
    short = np.asarray(short)
    long_ = np.asarray(long_)
    n, m = len(short), len(long_)
    best, frame = -1, -1
    for i in range(m - n + 1):
        s = cosine_similarity(short, long_[i:i + n])
        best = max(best, s)
        if s == best:
            frame = i
    return best, frame


def vector_distance(hsv1, hsv2):
    """computeVectorDistance.py:22-43 -> (true cosine [[c]], quirk row, sum |a-b|)."""
    hsv1 = np.asarray(hsv1, dtype=float).reshape(-1, 1)
    hsv2 = np.asarray(hsv2, dtype=float).reshape(-1, 1)
    quirk = np.dot(hsv1, hsv2.T) / (np.linalg.norm(hsv1, axis=1) * np.linalg.norm(hsv2, axis=1))
    a = hsv1.reshape(-1)
    b = hsv2.reshape(-1)
    cos = np.array([[np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b))]])
    n = min(len(hsv1), len(hsv2))
    dist = float(np.abs(hsv1[:n, 0] - hsv2[:n, 0]).sum())
    return cos, quirk[0], dist
