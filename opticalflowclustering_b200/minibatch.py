"""MiniBatchKMeans on the GPU: the clustering behind the reference's colour quantisation script

    clt = MiniBatchKMeans(n_clusters = args["clusters"]); labels = clt.fit_predict(image)
    quant = clt.cluster_centers_.astype("uint8")[labels]            (color-quantization/quant.py:18-20)

following scikit-learn 1.9.0's ``MiniBatchKMeans.fit`` step for step (sklearn/cluster/_kmeans.py:2056-2227,
``_mini_batch_step`` :1566-1684, ``_mini_batch_convergence`` :1974-2037, ``_random_reassign`` :2039-2054,
``_minibatch_update_dense`` in _k_means_minibatch.pyx) including the order of its ``RandomState`` calls, so that
``MiniBatchKMeans(n_clusters=k, random_state=int)`` reproduces scikit-learn on uint8 rows (the script's LAB pixels):
same mini-batches, same seeds, same centres, same labels.

Where the work runs: the rows stay on the device; per step the mini-batch is gathered on the device, labelled by the
E-step kernel (``ofc_kmeans_assign``: ||c||^2 - 2 x.c, first strict minimum) and folded into the centres by
``ofc_minibatch_update`` in the reference's accumulation order.  The host keeps sklearn's bookkeeping -- the random
stream (numpy ``RandomState``, as sklearn), the smoothed-inertia early stopping and the low-count reassignment -- and
reads back k + 1 numbers per step.  torch is used for device memory only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import kmeans as _km
from .kmeans import LloydState, _as_batch, _Ctx, _ptr, _target_device


class MiniBatchKMeans:
    """The subset of ``sklearn.cluster.MiniBatchKMeans`` the reference uses: constructor keywords, ``fit``,
    ``fit_predict``, ``predict``, ``cluster_centers_``, ``labels_``, ``inertia_``, ``n_steps_``, ``n_iter_``."""

    def __init__(self, n_clusters=8, *, init="k-means++", max_iter=100, batch_size=1024, compute_labels=True,
                 random_state=None, tol=0.0, max_no_improvement=10, init_size=None, n_init="auto", reassignment_ratio=0.01):
        self.n_clusters, self.init, self.max_iter, self.batch_size = int(n_clusters), init, int(max_iter), int(batch_size)
        self.compute_labels, self.random_state, self.tol = bool(compute_labels), random_state, float(tol)
        self.max_no_improvement, self.init_size, self.n_init = max_no_improvement, init_size, n_init
        self.reassignment_ratio = float(reassignment_ratio)
        if self.reassignment_ratio < 0:
            raise ValueError(f"reassignment_ratio should be >= 0, got {self.reassignment_ratio} instead.")

    # -- helpers ---------------------------------------------------------------------------------------------
    def _labels_inertia(self, st: LloydState, centres: torch.Tensor, labels: torch.Tensor):
        st.assign(None, centres.view(1, self.n_clusters, -1), labels, inertia=st.inertia)
        return st.inertia

    def fit(self, X, y=None, sample_weight=None):
        if sample_weight is not None:
            raise NotImplementedError("sample_weight is not used by the reference (color-quantization/quant.py:18-19)")
        Xb, single = _as_batch(X, _target_device(X))
        if not single:
            raise ValueError("Expected 2D array")
        ctx = _Ctx(Xb.device)
        dev = Xb.device
        Xf = Xb[0]
        n, d = int(Xf.shape[0]), int(Xf.shape[1])
        k = self.n_clusters
        if n < k:
            raise ValueError(f"n_samples={n} should be >= n_clusters={k}.")
        is_f32 = Xf.dtype == torch.float32
        # _check_params_vs_input (_kmeans.py:1932-1962)
        bs = min(self.batch_size, n)
        init_size = self.init_size
        if init_size is None:
            init_size = 3 * bs
            if init_size < k:
                init_size = 3 * k
        elif init_size < k:
            init_size = 3 * k
        init_size = min(init_size, n)
        if isinstance(self.init, str):
            if self.init != "k-means++":
                raise NotImplementedError(f"init={self.init!r}")
            n_init = 1 if self.n_init == "auto" else int(self.n_init)
        else:
            n_init = 1
        tol_ = 0.0
        if self.tol != 0:
            mean, var, _ = _km.column_mean_var(LloydState(ctx, Xb, 1))
            tol_ = float(var.mean().item()) * self.tol
        rs = self.random_state if isinstance(self.random_state, np.random.RandomState) else np.random.RandomState(self.random_state)

        # validation set for the init (:2110-2112); drawn even when a single init makes it unused
        valid_idx = torch.from_numpy(rs.randint(0, n, init_size)).to(dev)
        best = None
        for _ in range(n_init):
            if isinstance(self.init, str):
                Xi = Xf
                if init_size < n:                                       # _init_centroids (:1012-1017)
                    init_idx = torch.from_numpy(rs.randint(0, n, init_size)).to(dev)
                    Xi = Xf.index_select(0, init_idx).contiguous()
                centres, _ = _km.kmeans_plusplus(Xi, k, random_state=rs)
            else:
                centres = torch.as_tensor(np.asarray(self.init), dtype=torch.float64).to(dev).clone()
            if is_f32:
                centres = centres.to(torch.float32).to(torch.float64)
            if n_init > 1:
                Xv = Xf.index_select(0, valid_idx).contiguous().unsqueeze(0)
                stv = LloydState(ctx, Xv, k)
                inertia = float(self._labels_inertia(stv, centres, stv.labels[0]).item())
                if best is None or inertia < best[0]:
                    best = (inertia, centres)
            else:
                best = (0.0, centres)
        centres = best[1].contiguous()
        centres_new = torch.empty_like(centres)
        counts = torch.zeros(k, dtype=torch.float64, device=dev)            # self._counts (X.dtype in sklearn; float32 values stay exact)

        # the weighted sampling of the mini-batches (:2165-2170) is RandomState.choice(n, bs, p=1/n): a cumulative table
        # (the same every step) searched with `bs` uniform draws -- the table is built once, the draws are numpy's
        p = np.ones(n, dtype=np.float64) / float(n)
        cdf = p.cumsum()
        cdf /= cdf[-1]

        Xmb = torch.empty((1, bs, d), dtype=Xf.dtype, device=dev)
        st = LloydState(ctx, Xmb, k)
        labels_mb = st.labels[0]
        lib = ctx.lib
        ewa, ewa_min, no_improvement, n_since_reassign = None, None, 0, 0
        counts_h = np.zeros(k)
        n_steps = (self.max_iter * n) // bs
        step = -1
        for step in range(n_steps):
            idx = cdf.searchsorted(rs.random_sample(bs), side="right")
            torch.index_select(Xf, 0, torch.from_numpy(np.asarray(idx, dtype=np.int64)).to(dev), out=Xmb[0])
            # _random_reassign (:2039-2054), evaluated on the counts before this step
            n_since_reassign += bs
            random_reassign = bool((counts_h == 0).any()) or n_since_reassign >= 10 * k
            if random_reassign:
                n_since_reassign = 0
            # _mini_batch_step: labels + inertia with the current centres, then the centre update
            inertia_t = self._labels_inertia(st, centres, labels_mb)
            ctx.check(lib.ofc_minibatch_update(_ptr(Xmb), st.dtype, bs, d, k, _ptr(labels_mb), _ptr(centres), _ptr(centres_new),
                                               _ptr(counts), ctx.stream()))
            host = torch.cat([inertia_t.view(1), counts]).cpu().numpy()
            batch_inertia, counts_h = float(host[0]), host[1:].copy()
            if random_reassign and self.reassignment_ratio > 0:
                to_reassign = counts_h < self.reassignment_ratio * counts_h.max()
                if to_reassign.sum() > 0.5 * bs:
                    keep = np.argsort(counts_h)[int(0.5 * bs):]
                    to_reassign[keep] = False
                n_re = int(to_reassign.sum())
                if n_re:
                    new_rows = rs.choice(bs, replace=False, size=n_re)
                    rows = Xmb[0].index_select(0, torch.from_numpy(np.asarray(new_rows, dtype=np.int64)).to(dev)).to(torch.float64)
                    centres_new[torch.from_numpy(np.nonzero(to_reassign)[0]).to(dev)] = rows
                counts_h[to_reassign] = np.min(counts_h[~to_reassign])
                counts.copy_(torch.from_numpy(counts_h))
            sq_diff = float(((centres_new - centres) ** 2).sum().item()) if tol_ > 0.0 else 0.0
            centres, centres_new = centres_new, centres
            # _mini_batch_convergence (:1974-2037)
            batch_inertia /= bs
            if step + 1 == 1:
                continue
            if ewa is None:
                ewa = batch_inertia
            else:
                alpha = min(bs * 2.0 / (n + 1), 1)
                ewa = ewa * (1 - alpha) + batch_inertia * alpha
            if tol_ > 0.0 and sq_diff <= tol_:
                break
            if ewa_min is None or ewa < ewa_min:
                no_improvement, ewa_min = 0, ewa
            else:
                no_improvement += 1
            if self.max_no_improvement is not None and no_improvement >= self.max_no_improvement:
                break
        self._centres_t = centres.clone()
        self.cluster_centers_ = centres.cpu().numpy()
        if is_f32:
            self.cluster_centers_ = self.cluster_centers_.astype(np.float32)
        self.n_steps_ = step + 1
        self.n_iter_ = int(np.ceil(((step + 1) * bs) / n))
        if self.compute_labels:
            stf = LloydState(ctx, Xb, k)
            inertia = self._labels_inertia(stf, centres, stf.labels[0])
            self._labels_t = stf.labels[0][0]
            self.labels_ = self._labels_t.cpu().numpy()
            self.inertia_ = float(inertia.item())
        else:
            self.inertia_ = float(ewa * n) if ewa is not None else 0.0
        return self

    def fit_predict(self, X, y=None, sample_weight=None):
        return self.fit(X, sample_weight=sample_weight).labels_

    def predict(self, X):
        return _km.predict(X, self._centres_t).cpu().numpy()


def quantize(image_lab, n_clusters, random_state=None):
    """quant.py:15-20 on an image that is already in LAB (uint8 [H, W, 3], numpy or CUDA tensor): returns
    ``(labels [H*W], quantised image uint8 [H, W, 3])``; ``cluster_centers_.astype('uint8')[labels]`` is gathered on the
    device."""
    h, w = int(image_lab.shape[0]), int(image_lab.shape[1])
    flat = image_lab.reshape(h * w, 3)
    clt = MiniBatchKMeans(n_clusters=n_clusters, random_state=random_state)
    clt.fit(flat)
    cen_u8 = clt._centres_t.to(torch.uint8)                       # .astype("uint8"): truncation, values are in 0..255
    quant = cen_u8.index_select(0, clt._labels_t.long()).reshape(h, w, 3)
    if isinstance(image_lab, torch.Tensor):
        return clt._labels_t, quant
    return clt.labels_, quant.cpu().numpy()
