"""CPU baseline: the reference's hot path as it ships, i.e. Python glue around
cv2 4.13 and scikit-learn 1.9 (the wheels the reference imports).

TEST / BASELINE INFRASTRUCTURE ONLY: imported by tests/ and by bench.py's
``cpu_baseline`` and ``--impl reference`` legs, never by the product package.

/root/reference does not travel to the GPU box and has no installable package
(no setup.py / pyproject), so this module restates the reference's own call
sequence, call for call, on top of the same third-party libraries:
  FlowState.compute          computeOpticalFlowModule.py:18-36
  grid_pass                  KmeanGrids.py:52-113 (overlayGridAndComputeAvgColor)
  preprocess / cluster_hue   KmeanGrids.py:269-339 (KMeans(n_clusters=k) per cell)
  run_frames                 KmeanGrids.py:180-231 + :376-399 (main loops)
tests/test_reference_chain.py checks it against the golden outputs generated
from the real reference modules (tests/golden/make_golden.py).
"""
from __future__ import annotations

import os
import time

import numpy as np


def _cv2():
    import cv2
    return cv2


class FlowState:
    def __init__(self, first_frame):
        cv2 = _cv2()
        self.mask = np.zeros_like(first_frame)
        self.mask[..., 1] = 255
        self.prev_gray = cv2.cvtColor(first_frame, cv2.COLOR_BGR2GRAY)

    def compute(self, frame):
        cv2 = _cv2()
        gray = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)
        flow = cv2.calcOpticalFlowFarneback(self.prev_gray, gray, None, 0.5, 3, 15, 3, 5, 1.2, 0)
        magnitude, angle = cv2.cartToPolar(flow[..., 0], flow[..., 1])
        self.mask[..., 0] = angle * 180 / np.pi / 2
        self.mask[..., 2] = cv2.normalize(magnitude, None, 0, 255, cv2.NORM_MINMAX)
        self.prev_gray = gray
        return cv2.cvtColor(self.mask, cv2.COLOR_HSV2BGR)


def grid_pass(frame, rows=14, cols=25):
    """Per-cell mean -> uint8 -> hue, white rectangle, ROI views (reference order)."""
    cv2 = _cv2()
    h, w = frame.shape[:2]
    xs, ys = int(w / cols), int(h / rows)
    hues, rois = [], []
    for y in range(rows):
        for x in range(cols):
            x1, y1 = x * xs, y * ys
            x2, y2 = min(x1 + xs, w), min(y1 + ys, h)
            roi = frame[y1:y2, x1:x2]
            cv2.cvtColor(roi, cv2.COLOR_BGR2HSV)                       # computed and dropped, as in the reference
            avg = np.mean(roi, axis=(0, 1)).astype(np.uint8)
            hues.append(cv2.cvtColor(np.array([[avg]]), cv2.COLOR_BGR2HSV)[0, 0][0])
            cv2.rectangle(frame, (x1, y1), (x2, y2), (255, 255, 255), 1)
            rois.append(roi)
    return np.array(hues), rois


def preprocess(image):
    cv2 = _cv2()
    image[image < 30] = 0
    gray = cv2.cvtColor(image, cv2.COLOR_BGR2GRAY)
    _, alpha = cv2.threshold(gray, 0, 255, cv2.THRESH_BINARY)
    alpha[alpha > 0] = 255
    b, g, r = cv2.split(image)
    return cv2.merge([b, g, r, alpha], 4)


def cluster_hue(image4, n_clusters=1):
    cv2 = _cv2()
    from sklearn.cluster import KMeans
    flat = image4.reshape(image4.shape[0] * image4.shape[1], 4)
    clt = KMeans(n_clusters=n_clusters)
    clt.fit(flat)
    labels = clt.predict(flat)
    share = np.bincount(labels).astype(float) / len(flat)
    info = sorted([(share[i], i, c) for i, c in enumerate(clt.cluster_centers_)], key=lambda t: t[0], reverse=True)
    c = np.rint(info[0][2])
    hsv = cv2.cvtColor(np.array([[[c[0], c[1], c[2]]]], dtype=np.uint8), cv2.COLOR_BGR2HSV)
    return c, hsv[0][0][0]


def run_frames(frames, n_clusters=1, rows=14, cols=25):
    """frames u8[T,H,W,3] -> (avg_hue[T-1,cells], km_hue[T-1,cells]) like KmeanGrids.py's two loops."""
    state = FlowState(frames[0])
    avg_rows, km_rows = [], []
    for t in range(1, len(frames)):
        viz = state.compute(frames[t])
        hues, rois = grid_pass(viz, rows, cols)
        avg_rows.append(hues)
        km_rows.append([cluster_hue(preprocess(r), n_clusters)[1] for r in rois])
    return np.array(avg_rows), np.array(km_rows)


def _worker(args):
    frames, n_clusters, rows, cols = args
    import warnings
    warnings.filterwarnings("ignore")
    try:
        _cv2().setNumThreads(1)
    except Exception:
        pass
    run_frames(frames[:2], n_clusters, 2, 2)          # import / first-call costs outside the timed part
    t0 = time.perf_counter()
    run_frames(frames, n_clusters, rows, cols)
    return time.perf_counter() - t0


def timed_throughput(frames, workers: int, pairs_per_worker: int, n_clusters=1, rows=14, cols=25):
    """Frame pairs / s of the reference chain with ``workers`` processes, each
    running ``pairs_per_worker`` consecutive pairs (OpenCV's CPU Farneback does not
    scale with threads, so the many-core figure is process-parallel over frame
    ranges).  Workers are *spawned* (forking a process that holds
    torch / OpenMP threads deadlocks) and time only their own compute; the
    throughput is total pairs / slowest worker.  Returns (pairs_per_second, seconds)."""
    import multiprocessing as mp
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    T = len(frames)
    jobs = []
    for w in range(workers):
        s = (w * pairs_per_worker) % max(T - pairs_per_worker, 1)
        jobs.append((np.ascontiguousarray(frames[s:s + pairs_per_worker + 1]), n_clusters, rows, cols))
    if workers == 1:
        times = [_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(workers) as pool:
            times = pool.map(_worker, jobs, chunksize=1)
    slowest = max(times)
    return workers * pairs_per_worker / slowest, slowest
