"""Pin the grid / k-means / cosine oracle against the reference's goldens G1-G6."""
import csv
import os

import numpy as np
import pytest

from oracle import grid_np as G
from oracle import kmeans_np as K
from oracle import viz_np as V
from tests.conftest import GOLDEN, unpack_images


def _hue_col(path):
    with open(path, encoding="utf-8-sig") as f:
        return np.array([int(r[1]) for r in csv.reader(f) if r])


def test_g1_cluster_centers_csv_text():
    z = np.load(os.path.join(GOLDEN, "g1_images.npz"))
    rows = list(csv.reader(open(os.path.join(GOLDEN, "g1_cluster_centers.csv"), newline="")))
    assert rows[0] == ["File name", "Cluster 1", "HSV Cluster 1", "Hue 0"]
    for im, name, row in zip(unpack_images(z), z["names"], rows[1:]):
        rgb = im[..., ::-1].copy()                       # read_image: BGR -> RGB (Q5)
        c, h = G.cluster_colors_k1(G.preprocess_image(rgb))
        hsv0 = V.bgr2hsv_u8(np.array([[[c[0], c[1], c[2]]]], dtype=np.uint8))
        assert [str(name), str(c), str(hsv0), str(h)] == row


def test_g2_outcsv_hues():
    z = np.load(os.path.join(GOLDEN, "g23_cells.npz"))
    for fi in range(z["cells"].shape[0]):
        for c in range(350):
            rgb = z["cells"][fi, c][..., ::-1].copy()
            _, h = G.cluster_colors_k1(G.preprocess_image(rgb))
            assert h == z["outcsv_hues"][fi, c]


def test_g3_rgb_values_interior_cells():
    """cells saved after their own rectangle: interior cells (cx>0, cy>0) had both
    white lines at mean time, so mean(saved cell) reproduces the CSV (Q3)."""
    z = np.load(os.path.join(GOLDEN, "g23_cells.npz"))
    for fi in range(z["cells"].shape[0]):
        for c in range(350):
            cy, cx = divmod(c, 25)
            if cy == 0 or cx == 0:
                continue
            roi = z["cells"][fi, c]
            n = roi.shape[0] * roi.shape[1]
            avg = (roi.reshape(n, 3).astype(np.int64).sum(0) // n).astype(np.uint8)
            assert V.bgr2hsv_u8(avg[None, None])[0, 0, 0] == int(z["rgb_values_hues"][fi, c])


def test_g4_hues():
    z = np.load(os.path.join(GOLDEN, "g4_images.npz"))
    for im, want in zip(unpack_images(z), z["hues"]):
        _, h = G.cluster_colors_k1(G.preprocess_image(im[..., ::-1].copy()))
        assert h == want


def test_g5_sliding_cosine():
    a = _hue_col(os.path.join(GOLDEN, "bounce.csv"))
    b = _hue_col(os.path.join(GOLDEN, "601_3_3_cropped.csv"))
    assert (len(a), len(b)) == (16, 75)
    s, f = G.sliding_cosine(a, b)
    assert repr(float(s)) == "0.91448231723348" and f == 24
    s, f = G.sliding_cosine(a, _hue_col(os.path.join(GOLDEN, "cropped_trimmed2.csv")))
    assert repr(float(s)) == "0.963475622684391" and f == 7


def test_g6_vector_distance():
    """computeVectorDistance.py on the reference's file1.csv / file2.csv prints
    [[1.]], the quirk row below and distance 0.0 (run of the reference script)."""
    a = _hue_col(os.path.join(GOLDEN, "file1.csv"))
    b = _hue_col(os.path.join(GOLDEN, "file2.csv"))
    cos, quirk, dist = G.vector_distance(a, b)
    assert str(cos) == "[[1.]]"
    want = [1., 1.09756098, 0.91836735, 0.83333333, 0.8490566, 0.52023121, 0.62068966, 0.52631579, 0.55900621,
            0.5625, 0.58441558, 1.02272727, 1.04651163, 1.23287671, 1.34328358, 1.52542373]
    assert np.allclose(quirk, want, rtol=0, atol=5e-9)
    assert dist == 0.0


@pytest.mark.parametrize("name", ["u8_d4_k8", "f32_d32_k16", "u8_d4_k3"])
def test_kmeans_vs_sklearn_golden(name):
    z = np.load(os.path.join(GOLDEN, "kmeans_sklearn.npz"))
    labels, centers, inertia, n_iter = K.kmeans_fit(z[name + "_X"], z[name + "_init"])
    assert (labels == z[name + "_labels"]).all()
    assert n_iter == int(z[name + "_niter"])
    tol = 1e-6 if name.startswith("f32") else 1e-12
    assert abs(inertia - float(z[name + "_inertia"])) <= tol * float(z[name + "_inertia"])
    assert np.allclose(centers, z[name + "_centers"], rtol=0, atol=1e-4 if name.startswith("f32") else 1e-9)
    assert (K.kmeans_predict(z[name + "_X"], centers) == z[name + "_predict"]).all()


def test_rint_mean_exact_matches_sklearn_k1():
    pytest.importorskip("sklearn")
    from sklearn.cluster import KMeans
    rng = np.random.default_rng(3)
    for n in (2500, 2601, 50 * 52):
        X = rng.integers(0, 256, (n, 4)).astype(np.uint8)
        X[:, 3] = np.where(rng.random(n) < 0.5, 255, 0)
        X[: n // 2, 0] = 7
        X[n // 2:, 0] = 8                               # exact .5 tie in column 0 for even n
        km = KMeans(n_clusters=1).fit(X)
        want = np.rint(km.cluster_centers_[0])
        got = K.rint_mean_exact(X.astype(np.int64).sum(0), n)
        assert (want == got).all(), (want, got)


def test_oracle_kmeans_sweep_ends_vs_sklearn_golden():
    """the oracle against the committed sklearn results at D = 2 / 2048 / 2050 / 5000 (tests/golden/make_golden.py:
    make_sweep_kmeans_golden): labels bit-exact, same n_iter"""
    import os
    import sys
    import numpy as np
    from oracle import kmeans_np as K
    from tests.conftest import GOLDEN
    sys.path.insert(0, GOLDEN)
    import make_golden as MG
    z = np.load(os.path.join(GOLDEN, "kmeans_sklearn_sweep.npz"))
    for name in MG.SWEEP_CASES:
        X, init = MG.sweep_case(name)
        w = K.kmeans_fit(X, init)
        assert (w[0] == z[name + "_labels"]).all() and w[3] == int(z[name + "_niter"]), name
        assert abs(w[2] - float(z[name + "_inertia"])) <= 1e-5 * float(z[name + "_inertia"])
