"""CPU oracle: numpy restatement of scikit-learn 1.9.0 KMeans (lloyd, dense).

TEST INFRASTRUCTURE ONLY (see oracle/farneback_np.py header for the rule).

The reference calls ``KMeans(n_clusters=k).fit(X)`` then ``.predict(X)``
(k-means-color-clustering/KmeanGrids.py:299-304, color_kmeans.py:65-78).  The
arithmetic is in the third-party wheel scikit-learn 1.9.0 (not vendored, not
pinned by the reference); its Cython sources are readable on this image:
  sklearn/cluster/_kmeans.py:1463-1546     fit(): dtype promotion, tol, centring
  sklearn/cluster/_kmeans.py:627-758       _kmeans_single_lloyd loop / stopping
  sklearn/cluster/_k_means_lloyd.pyx:160-213  E-step: ||c||^2 - 2 x.c, first strict min
  sklearn/cluster/_k_means_common.pyx:167-311 relocate empty, average, shift
This file restates that algorithm.

Parity pin: tests/test_oracle_kmeans.py against live sklearn with
``KMeans(init=C0, n_init=1)`` (k>1 is unpinned in the reference itself because
it never fixes random_state) and against the reference's k=1 goldens G1/G2/G4
(tests/golden/).
"""
from __future__ import annotations

import numpy as np


def _work_dtype(X: np.ndarray):
    return np.float32 if X.dtype == np.float32 else np.float64


def _e_step(X, centers):
    """labels = first strict minimum of ||c||^2 - 2 x.c  (ties -> lowest index)."""
    c2 = (centers * centers).sum(axis=1)
    n = X.shape[0]
    labels = np.empty(n, np.int32)
    chunk = 1 << 16
    for s in range(0, n, chunk):
        d = c2[None, :] - 2.0 * (X[s:s + chunk] @ centers.T)
        labels[s:s + chunk] = np.argmin(d, axis=1)
    return labels


def _m_step(X, labels, centers_old):
    k, D = centers_old.shape
    w = np.bincount(labels, minlength=k).astype(X.dtype)
    sums = np.zeros((k, D), X.dtype)
    for d in range(D):
        sums[:, d] = np.bincount(labels, weights=X[:, d], minlength=k)
    # relocate empty clusters to the farthest points (_k_means_common.pyx:167-211)
    empty = np.where(w == 0)[0]
    if empty.size:
        dist = ((X - centers_old[labels]) ** 2).sum(axis=1)
        if dist.max() > 0:
            far = np.argpartition(dist, -empty.size)[:-empty.size - 1:-1]
            for idx, new_id in enumerate(empty):
                fi = far[idx]
                old_id = labels[fi]
                sums[old_id] -= X[fi]
                sums[new_id] = X[fi]
                w[new_id] = 1
                w[old_id] -= 1
    new = np.empty_like(sums)
    amax = int(np.argmax(w))
    for j in range(k):
        if w[j] > 0:
            new[j] = sums[j] * (X.dtype.type(1.0) / w[j])
    for j in range(k):
        if not w[j] > 0:
            new[j] = new[amax]
    shift = np.sqrt(((new - centers_old) ** 2).sum(axis=1))
    return new, w, shift


def kmeans_fit(X, init, max_iter: int = 300, tol: float = 1e-4):
    """KMeans(n_clusters=k, init=init, n_init=1, max_iter, tol).fit(X).

    Returns ``(labels int32[N], centers[k,D], inertia, n_iter)`` with the
    semantics of ``labels_``, ``cluster_centers_``, ``inertia_``, ``n_iter_``.
    """
    X = np.asarray(X)
    dt = _work_dtype(X)
    X = X.astype(dt)                                      # uint8 -> float64
    centers = np.asarray(init).astype(dt).copy()
    tol_ = float(np.mean(np.var(X, axis=0)) * tol)        # _tolerance()
    mean = X.mean(axis=0)
    Xc = X - mean
    centers -= mean
    labels_old = np.full(X.shape[0], -1, np.int32)
    strict = False
    n_iter = 0
    labels = labels_old
    for it in range(max_iter):
        labels = _e_step(Xc, centers)
        new, w, shift = _m_step(Xc, labels, centers)
        centers = new
        n_iter = it + 1
        if np.array_equal(labels, labels_old):
            strict = True
            break
        if float((shift ** 2).sum()) <= tol_:
            break
        labels_old = labels
    if not strict:
        labels = _e_step(Xc, centers)
    diff = Xc - centers[labels]
    inertia = float((diff * diff).sum())
    return labels, (centers + mean), inertia, n_iter


def kmeans_predict(X, centers):
    """KMeans.predict: E-step on the un-centred data (_kmeans.py:1075-1107)."""
    X = np.asarray(X)
    dt = _work_dtype(X)
    return _e_step(X.astype(dt), np.asarray(centers).astype(dt))


def rint_mean_exact(sum_int: np.ndarray, n: int) -> np.ndarray:
    """np.rint(sum/n) in exact integer arithmetic (round-half-even).

    For k = 1 the fitted centre is the column mean; sklearn's centred fp64
    computation is within ~1e-14 of sum/n and exact when sum/n is a half
    integer, so np.rint of it equals this integer rule (checked
    against sklearn in tests/test_oracle_kmeans.py).
    """
    s = np.asarray(sum_int, dtype=np.int64)
    q, r = np.divmod(s, n)
    up = (2 * r > n) | ((2 * r == n) & (q % 2 == 1))
    return q + up.astype(np.int64)
