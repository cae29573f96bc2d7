"""Drop-in for the reference's k-means-color-clustering/computeVectorDistance.py
(byte-identical to exampleVectorDistances.py): hard-coded ``file1.csv`` / ``file2.csv`` in
the cwd (:6-7), columns 1.. of each row as floats, then the three printed results (:27,43-44).
The arithmetic is libofc's vector-distance kernel.
"""
from __future__ import annotations

import csv

import numpy as np

from .cosine import vector_distance


def read_rows(path):
    with open(path, 'r') as f:
        return np.array([row[1:] for row in csv.reader(f)], dtype=float)


def main(file1='file1.csv', file2='file2.csv'):
    hsv1, hsv2 = read_rows(file1), read_rows(file2)
    if hsv1.shape[1] != 1 or hsv2.shape[1] != 1:
        raise NotImplementedError("only the reference's two-column (name, hue) CSV layout is supported")
    similarity, cos_similarity, euclidean_distance = vector_distance(hsv1[:, 0], hsv2[:, 0])
    print(similarity)
    if len(hsv1) != len(hsv2):
        print("Warning: The vectors have different lengths, only the Euclidean distance of the common "
              "subvectors has been computed.")
    print("Cosine similarity:", cos_similarity)
    print("Euclidean distance:", euclidean_distance)
    return similarity, cos_similarity, euclidean_distance


if __name__ == "__main__":
    main()
