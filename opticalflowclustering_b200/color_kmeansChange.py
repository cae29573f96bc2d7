"""Drop-in for the reference's k-means-color-clustering/color_kmeansChange.py: color_kmeans.py run over
a directory tree ``<dir>/<frame>/<cell>.png`` (:149-159), one CSV row per image keyed by
``<frame>/<cell>.png`` instead of the basename (:133).

    python -m opticalflowclustering_b200.color_kmeansChange -d OutImgs/<video>/ -c 1 -f out.csv

For the documented ``-c 1`` every frame folder is clustered in ONE batched launch sequence (all its
images of equal size stacked as a [B, n, 4] problem for the Lloyd kernels); other k take the images
one by one through :func:`color_kmeans.dominant_cluster`.
"""
from __future__ import annotations

import argparse
import csv
import os
import re

import numpy as np

from . import color_kmeans as _ck
from . import kmeans as _km

read_image = _ck.read_image
preprocess_image = _ck.preprocess_image


def parse_arguments(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("-d", "--dir", required=True, help="Path to the image")
    ap.add_argument("-c", "--clusters", required=True, type=int, help="# of clusters")
    ap.add_argument("-f", "--csv", required=True, type=str, help="# of clusters")
    return vars(ap.parse_args(argv))


def get_number(filename):
    match = re.compile(r'(\d+)').search(filename)
    return int(match.group(1)) if match else None


def _write_row(csv_file, image_path, centre_rint):
    r0, g0, b0, _a0 = centre_rint
    hsv0 = _ck.bgr2hsv_pixels(np.array([[[r0, g0, b0]]], dtype=np.uint8))
    with open(csv_file, 'a', newline='') as file:
        writer = csv.writer(file)
        if os.stat('cluster_centers.csv').st_size == 0:          # hard-coded name, like the reference (:105)
            writer.writerow(_ck.HEADER)
        writer.writerow([image_path, centre_rint, hsv0, hsv0[0][0][0]])


def cluster_colors(image, n_clusters, image_path, csv_file, init=None, random_state=None):
    """color_kmeansChange.py:54-135: like color_kmeans.cluster_colors but the row is keyed by ``image_path``."""
    info, _ = _ck.dominant_cluster(np.asarray(image), n_clusters, init, random_state)
    _write_row(csv_file, image_path, np.rint(info[0][2]))
    return None


def cluster_folder_k1(images, keys, csv_file):
    """k = 1 for a whole frame folder: the fitted centre is the column mean, so all images of one size
    are one batched Lloyd problem."""
    by_shape = {}
    for i, im in enumerate(images):
        by_shape.setdefault(im.shape, []).append(i)
    centres = [None] * len(images)
    for shape, idx in by_shape.items():
        X = np.stack([images[i].reshape(-1, 4) for i in idx])
        _, c, _, _ = _km.lloyd(X, X[:, :1].astype(np.float64))
        c = c.cpu().numpy()
        for j, i in enumerate(idx):
            centres[i] = np.rint(c[j, 0])
    for key, c in zip(keys, centres):
        _write_row(csv_file, key, c)


def main(argv=None):
    args = parse_arguments(argv)
    dirs = args['dir']
    for contentFolder in sorted(os.listdir(dirs), key=get_number):
        names = sorted(os.listdir(dirs + contentFolder), key=get_number)
        images = [preprocess_image(read_image(dirs + contentFolder + '/' + n)) for n in names]
        keys = [contentFolder + '/' + n for n in names]
        if args["clusters"] == 1:
            cluster_folder_k1(images, keys, args["csv"])
        else:
            for im, key in zip(images, keys):
                cluster_colors(im, args["clusters"], key, args["csv"])
        print(contentFolder)


if __name__ == "__main__":
    main()
