"""Drop-in for the reference's k-means-color-clustering/color_kmeans.py (and the
``preprocess_image`` / ``cluster_colors`` pair of KmeanGrids.py:269-339).

Same function names, arguments, side effects and CSV text:

  read_image(path)                 cv2.imread + BGR->RGB                     (color_kmeans.py:28-33)
  preprocess_image(image)          in-place ``image[image < 30] = 0``; returns the 4-channel
                                   (c0, c1, c2, alpha) image                 (:35-52)
  cluster_colors(image, n_clusters, image_path, csv_file)
                                   KMeans -> predict -> bincount -> largest cluster ->
                                   np.rint -> BGR2HSV -> one CSV row; returns None   (:54-135)

All pixel arithmetic runs in libofc.so on the GPU (threshold/alpha in the cell-gather
kernel, k-means in the Lloyd kernels, BGR2HSV in the integer HSV kernel); cv2 is used for
image file decoding only.  numpy in -> numpy out, like the reference.
"""
from __future__ import annotations

import argparse
import csv
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from . import kmeans as _km
from .flow import _ptr, _stream_ptr, to_device_u8

THRESHOLD = 30          # color_kmeans.py:43, KmeanGrids.py:277
HEADER = ["File name", "Cluster 1", "HSV Cluster 1", "Hue 0"]      # color_kmeans.py:110


def parse_arguments(argv=None):
    """Same flags as color_kmeans.py:13-26 (-i image, -c clusters, -f csv)."""
    ap = argparse.ArgumentParser()
    ap.add_argument("-i", "--image", required=True, help="Path to the image")
    ap.add_argument("-c", "--clusters", required=True, type=int, help="# of clusters")
    ap.add_argument("-f", "--csv", required=True, type=str, help="# of clusters")
    return vars(ap.parse_args(argv))


def read_image(image_path):
    """cv2.imread then BGR -> RGB (color_kmeans.py:28-33); file decoding stays on the host."""
    import cv2
    image = cv2.imread(image_path)
    if image is None:
        raise cv2.error(f"could not read {image_path!r}")
    return np.ascontiguousarray(image[..., ::-1])


def preprocess_image_device(image_dev: torch.Tensor) -> torch.Tensor:
    """CUDA uint8 [H,W,3] -> CUDA uint8 [H,W,4]: channels below 30 zeroed, alpha = 255 where the
    BGR2GRAY of the thresholded pixel is > 0 (color_kmeans.py:43-52)."""
    H, W = int(image_dev.shape[0]), int(image_dev.shape[1])
    out = torch.empty((H, W, 4), dtype=torch.uint8, device=image_dev.device)
    with torch.cuda.device(image_dev.device):
        # the whole image is one "cell" of a 1x1 grid, no grid lines, no channel swap
        _lib.check(_lib.lib().ofc_grid_extract_cells(_ptr(image_dev), 1, H, W, 1, 1, 0, THRESHOLD, 0, _ptr(out),
                                                     _stream_ptr()))
    return out


def preprocess_image(image):
    """Same contract as the reference: mutates ``image`` in place (``image[image<30]=0``) and
    returns a new 4-channel array."""
    if isinstance(image, torch.Tensor):
        out = preprocess_image_device(image.contiguous())
        image.copy_(out[..., :3])
        return out
    if image.dtype != np.uint8 or image.ndim != 3 or image.shape[2] != 3:
        raise ValueError("expected a uint8 image [H, W, 3]")
    out = preprocess_image_device(to_device_u8(image)).cpu().numpy()
    image[...] = out[..., :3]
    return out


def bgr2hsv_pixels(bgr_u8: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(x, COLOR_BGR2HSV) for an array of 8-bit pixels [..., 3] (GPU integer kernel)."""
    a = np.ascontiguousarray(bgr_u8, dtype=np.uint8)
    dev = to_device_u8(a.reshape(-1, 3))
    out = torch.empty_like(dev)
    with torch.cuda.device(dev.device):
        _lib.check(_lib.lib().ofc_bgr2hsv(_ptr(dev), _ptr(out), C.c_int64(dev.shape[0]), _stream_ptr()))
    return out.cpu().numpy().reshape(a.shape)


def dominant_cluster(image, n_clusters, init=None, random_state=None):
    """fit -> predict -> bincount -> stable sort by share (descending) of color_kmeans.py:64-96.
    Returns ``(label_info, clt)`` with label_info = [(share, 'Cluster i', centre), ...] sorted."""
    flat = image.reshape(image.shape[0] * image.shape[1], 4)
    clt = _km.KMeans(n_clusters=n_clusters, init="k-means++" if init is None else init, random_state=random_state)
    clt.fit(flat)
    labels = clt.predict(flat)
    label_counts = np.bincount(labels)
    share = label_counts.astype(float) / len(flat)
    info = [(share[i], f"Cluster {i + 1}", c) for i, c in enumerate(clt.cluster_centers_)]
    info = sorted(info, key=lambda x: x[0], reverse=True)
    return info, clt


def cluster_row(image_path, centre_rint, hsv0):
    """The CSV row of color_kmeans.py:133 as a list of the objects csv.writer stringifies."""
    return [os.path.basename(image_path), centre_rint, hsv0, hsv0[0][0][0]]


def cluster_colors(image, n_clusters, image_path, csv_file, init=None, random_state=None):
    """color_kmeans.py:54-135: appends ``[basename, rint(centre), hsv, hue]`` to ``csv_file``
    (header iff the file named ``cluster_centers.csv`` in the cwd is empty -- the reference
    hard-codes that name, :107) and returns None."""
    print(image.shape)
    info, clt = dominant_cluster(np.asarray(image), n_clusters, init, random_state)
    print("Centroid clusters:", clt.cluster_centers_)
    print("Label", clt.labels_)
    for share, name, centre in info:
        print(f"{name}: {share * 100:.2f}%\nCluster Center: {np.rint(centre)}\n")
    with open(csv_file, "a", newline="") as file:
        writer = csv.writer(file)
        if os.stat("cluster_centers.csv").st_size == 0:
            writer.writerow(HEADER)
        c = np.rint(info[0][2])
        r0, g0, b0, _a0 = c
        rgb0 = np.array([[[r0, g0, b0]]], dtype=np.uint8)
        hsv0 = bgr2hsv_pixels(rgb0)                       # BGR2HSV applied to the RGB-ordered triple (Q5)
        print(rgb0)
        print("HSVs", hsv0[0][0])
        writer.writerow(cluster_row(image_path, c, hsv0))
    return None


def main(argv=None):
    args = parse_arguments(argv)
    image = read_image(args["image"])
    print("\n\n\n Image Name", args["image"])
    processed_image = preprocess_image(image)
    print("Dimensions", processed_image.ndim)
    cluster_colors(processed_image, args["clusters"], args["image"], args["csv"])


if __name__ == "__main__":
    main()
