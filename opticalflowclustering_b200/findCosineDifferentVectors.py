"""Drop-in for the reference's k-means-color-clustering/findCosineDifferentVectors.py.

    python -m opticalflowclustering_b200.findCosineDifferentVectors short.csv long.csv

Column 1 of two header-less CSVs (UTF-8 BOM tolerated, like pandas), the short vector slid
over the long one, same three output lines (:45,64-66).  The similarities, their maximum and
the LAST arg-max come from libofc's sliding-cosine kernels.
"""
from __future__ import annotations

import csv
import sys

import numpy as np

from .cosine import calculate_cosine_similarity, sliding_cosine  # noqa: F401  (same public name as the reference)


def read_hue_column(path):
    """``pd.read_csv(path, header=None).iloc[:, 1].values`` for the reference's hue CSVs."""
    with open(path, encoding="utf-8-sig", newline="") as f:
        col = [r[1] for r in csv.reader(f) if r]
    try:
        return np.array([int(v) for v in col], dtype=np.int64)
    except ValueError:
        return np.array([float(v) for v in col], dtype=np.float64)


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    file1_hue = read_hue_column(argv[0])
    nobounce_hue = read_hue_column(argv[1])
    smaller_len, larger_len = len(file1_hue), len(nobounce_hue)
    print("Vector sizes are: ", smaller_len, larger_len)
    max_similarity, max_frame = sliding_cosine(file1_hue, nobounce_hue)
    min_euclidean = 0                                   # constant in the reference (:50,65)
    print("Maximum cosine similarity:", max_similarity)
    print("Minimum sum of squared differences:", min_euclidean)
    print("Max frame:", max_frame)
    return max_similarity, max_frame


if __name__ == "__main__":
    main()
