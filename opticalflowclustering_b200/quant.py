"""Drop-in for the reference's color-quantization/quant.py: colour quantisation of an image by
MiniBatchKMeans in L*a*b* space.

    python -m opticalflowclustering_b200.quant -i image.png -c 8 [-o out.png] [--seed 0]

Same flags as the reference (``-i/--image``, ``-c/--clusters``; quant.py:6-10).  The clustering -- ``MiniBatchKMeans(
n_clusters).fit_predict`` and the ``cluster_centers_.astype("uint8")[labels]`` gather (quant.py:18-20) -- runs on the
GPU (minibatch.py).  The two ``cv2.cvtColor`` calls (BGR -> LAB before, LAB -> BGR after, quant.py:15,25-26) stay on the
host with cv2 like the image decode: OpenCV's 8-bit Lab conversion is a soft-float-built trilinear table, not
restated here.  The reference shows the result in a window (``cv2.imshow``); the headless drop-in returns the side-by-side
image and writes it when ``-o`` is given.
"""
from __future__ import annotations

import argparse

import numpy as np

from .minibatch import quantize


def parse_arguments(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("-i", "--image", required=True, help="Path to the image")
    ap.add_argument("-c", "--clusters", required=True, type=int, help="# of clusters")
    ap.add_argument("-o", "--output", default=None, help="write np.hstack([image, quant]) here (the reference shows it)")
    ap.add_argument("--seed", type=int, default=None, help="random_state of MiniBatchKMeans (the reference leaves it unset)")
    return vars(ap.parse_args(argv))


def quantize_image(image_bgr: np.ndarray, n_clusters: int, random_state=None):
    """quant.py:12-26 for a decoded BGR image: returns ``(image, quant)`` as BGR uint8 arrays"""
    import cv2
    lab = cv2.cvtColor(image_bgr, cv2.COLOR_BGR2LAB)
    _, quant = quantize(lab, n_clusters, random_state=random_state)
    return cv2.cvtColor(lab, cv2.COLOR_LAB2BGR), cv2.cvtColor(quant, cv2.COLOR_LAB2BGR)


def main(argv=None):
    import cv2
    args = parse_arguments(argv)
    image = cv2.imread(args["image"])
    if image is None:
        raise cv2.error(f"could not read {args['image']!r}")
    image, quant = quantize_image(image, args["clusters"], args["seed"])
    side = np.hstack([image, quant])
    if args["output"]:
        cv2.imwrite(args["output"], side)
    return side


if __name__ == "__main__":
    main()
