"""Generate tests/golden/* from the real reference (run in the build container only).

    python tests/golden/make_golden.py

Needs /root/reference (read-only) plus the cv2 / sklearn wheels the reference
calls.  The GPU box has no /root/reference, so everything the ``-m gpu`` tests
need is written here as small fixtures and committed.  What it does:

* imports the reference's own modules from
  /root/reference/k-means-color-clustering with three shims (stub matplotlib,
  no-op cv2.waitKey/imshow/destroyAllWindows, cwd with a cluster_centers.csv);
* G1  images/601_3_cropped_1_OF/cropped/*.png -> cluster_centers.csv: re-runs
  the reference's color_kmeans.cluster_colors and asserts the text is identical
  to the file the reference ships, then stores images + text;
* G2/G3 OutImgs/601_bad_bounce_3/{2,3,4}/*.png with the matching rows of
  OutCSV/601_bad_bounce_3.csv and 601_bad_bounce_3.mp4_rgb_values.csv;
* G4  images/cropped_trimmed_2/cropped/*.png -> cropped_trimmed2.csv;
* G5/G6 the small hue CSVs and the expected printed numbers;
* flow: cv2.calcOpticalFlowFarneback and the reference's
  ComputeOpticalFLow.compute on seeded synthetic clips;
* k-means: sklearn KMeans(init=C0, n_init=1) labels / centres / inertia.
"""
from __future__ import annotations

import csv
import io
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/k-means-color-clustering"
sys.path.insert(0, ROOT)


def import_reference():
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    import cv2
    cv2.waitKey = lambda *a, **k: -1
    cv2.imshow = lambda *a, **k: None
    cv2.destroyAllWindows = lambda *a, **k: None
    sys.path.insert(0, REF)
    import color_kmeans                       # noqa: E402  (reference module)
    import computeOpticalFlowModule           # noqa: E402
    import KmeanGrids                         # noqa: E402
    return cv2, color_kmeans, computeOpticalFlowModule, KmeanGrids


def pack_images(imgs):
    """ragged list of HxWx3 u8 -> (shapes int32[n,3], flat u8) for np.savez."""
    shapes = np.array([im.shape for im in imgs], np.int32)
    flat = np.concatenate([im.reshape(-1) for im in imgs]).astype(np.uint8)
    return shapes, flat


def main():
    cv2, color_kmeans, cofm, kmg = import_reference()
    from sklearn.cluster import KMeans
    import torch
    from opticalflowclustering_b200.synthetic import synthetic_clip

    # ---------------- G1 ----------------
    d = f"{REF}/images/601_3_cropped_1_OF/cropped"
    names = sorted(n for n in os.listdir(d) if n.endswith(".png"))
    imgs = [cv2.imread(f"{d}/{n}") for n in names]                    # BGR as on disk
    want = open(f"{REF}/cluster_centers.csv", newline="").read()
    with tempfile.TemporaryDirectory() as tmp:
        cwd = os.getcwd()
        os.chdir(tmp)
        open("cluster_centers.csv", "w").close()
        out = io.StringIO()
        so = sys.stdout
        sys.stdout = out
        try:
            for n in names:
                im = color_kmeans.read_image(f"{d}/{n}")
                pim = color_kmeans.preprocess_image(im)
                color_kmeans.cluster_colors(pim, 1, f"{d}/{n}", "cluster_centers.csv")
        finally:
            sys.stdout = so
            got = open("cluster_centers.csv", newline="").read()
            os.chdir(cwd)
    assert got == want, "reference no longer reproduces its own cluster_centers.csv"
    shapes, flat = pack_images(imgs)
    np.savez_compressed(f"{HERE}/g1_images.npz", shapes=shapes, flat=flat, names=np.array(names))
    open(f"{HERE}/g1_cluster_centers.csv", "w", newline="").write(want)
    print("G1 ok:", len(names), "images; csv text identical")

    # ---------------- G2 / G3 ----------------
    base = f"{REF}/OutImgs/601_bad_bounce_3"
    frames = [2, 3, 4]
    cells = np.stack([np.stack([cv2.imread(f"{base}/{f}/{c}.png") for c in range(1, 351)]) for f in frames])
    out_rows = list(csv.reader(open(f"{REF}/OutCSV/601_bad_bounce_3.csv")))
    rgb_rows = list(csv.reader(open(f"{REF}/601_bad_bounce_3.mp4_rgb_values.csv")))
    g2 = np.array([[int(v) for v in out_rows[1 + (f - 2)]] for f in frames])
    g3 = np.array([[float(v) for v in rgb_rows[1 + (f - 2)]] for f in frames])
    np.savez_compressed(f"{HERE}/g23_cells.npz", cells=cells, frames=np.array(frames),
                        outcsv_hues=g2, rgb_values_hues=g3)
    # confirm with the reference's own functions (G2): read_image -> preprocess -> KMeans(1)
    mism = 0
    for fi, f in enumerate(frames[:1]):
        for c in range(350):
            im = cv2.cvtColor(cells[fi, c].copy(), cv2.COLOR_BGR2RGB)
            pim = kmg.preprocess_image(im)
            _, hue = kmg.cluster_colors(pim, 1, "x", os.devnull)
            mism += int(hue != g2[fi, c])
    assert mism == 0, mism
    print("G2/G3 ok:", cells.shape)

    # ---------------- G4 ----------------
    d = f"{REF}/images/cropped_trimmed_2/cropped"
    names = sorted(n for n in os.listdir(d) if n.endswith(".png"))
    imgs = [cv2.imread(f"{d}/{n}") for n in names]
    rows = list(csv.reader(open(f"{REF}/cropped_trimmed2.csv", encoding="utf-8-sig")))
    hues = {r[0]: int(r[-1]) for r in rows if r and r[0].endswith(".png")}
    shapes, flat = pack_images(imgs)
    np.savez_compressed(f"{HERE}/g4_images.npz", shapes=shapes, flat=flat, names=np.array(names),
                        hues=np.array([hues[n] for n in names]))
    print("G4 ok:", len(names))

    # ---------------- G5 / G6 ----------------
    for n in ["bounce.csv", "601_3_3_cropped.csv", "cropped_trimmed2.csv", "file1.csv", "file2.csv"]:
        open(f"{HERE}/{n}", "wb").write(open(f"{REF}/{n}", "rb").read())
    print("G5/G6 copied")

    # ---------------- flow ----------------
    for (H, W, seed) in [(96, 128, 3), (135, 240, 4), (270, 480, 5)]:
        clip = synthetic_clip(3, H, W, seed=seed).numpy()
        gray = np.stack([cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in clip])
        flows = np.stack([cv2.calcOpticalFlowFarneback(gray[i], gray[i + 1], None, 0.5, 3, 15, 3, 5, 1.2, 0)
                          for i in range(2)])
        ref = cofm.ComputeOpticalFLow(clip[0].copy())
        viz = np.stack([ref.compute(clip[i + 1].copy()) for i in range(2)])
        np.savez_compressed(f"{HERE}/flow_{H}x{W}.npz", clip=clip, gray=gray, flow=flows, viz=viz)
        print("flow golden", H, W, "mean |flow|", np.abs(flows).mean())

    # ---------------- k-means ----------------
    rng = np.random.default_rng(7)
    cases = {}
    for name, (N, D, k, dtype) in {"u8_d4_k8": (20000, 4, 8, np.uint8), "f32_d32_k16": (8000, 32, 16, np.float32),
                                   "u8_d4_k3": (2601, 4, 3, np.uint8)}.items():
        cen = rng.uniform(20, 230, (k, D))
        X = cen[rng.integers(k, size=N)] + rng.normal(0, 12, (N, D))
        X = np.clip(np.rint(X), 0, 255).astype(np.uint8) if dtype == np.uint8 else X.astype(np.float32)
        init = X[:k].astype(np.float64 if dtype == np.uint8 else np.float32)
        km = KMeans(n_clusters=k, init=init, n_init=1).fit(X)
        cases[name + "_X"] = X
        cases[name + "_init"] = init
        cases[name + "_labels"] = km.labels_.astype(np.int32)
        cases[name + "_centers"] = km.cluster_centers_
        cases[name + "_inertia"] = np.float64(km.inertia_)
        cases[name + "_niter"] = np.int64(km.n_iter_)
        cases[name + "_predict"] = km.predict(X).astype(np.int32)
    np.savez_compressed(f"{HERE}/kmeans_sklearn.npz", **cases)
    print("kmeans goldens ok")
    make_dense_kmeans_golden()
    make_seeded_kmeans_golden()
    make_sweep_kmeans_golden()
    make_minibatch_golden()


def make_dense_kmeans_golden():
    """float32 rows with d > 32 (the tensor-core path of the build): sklearn 1.9.0 KMeans(init=first k rows, n_init=1).
    Separate file so the older fixtures stay byte-identical; X is regenerated from the seed by the test."""
    from sklearn.cluster import KMeans
    cases = {}
    for name, (N, D, k, seed) in {"f32_d64_k32": (6000, 64, 32, 11), "f32_d96_k40": (5000, 96, 40, 12)}.items():
        rng = np.random.default_rng(seed)
        cen = rng.uniform(0, 8, (k, D))
        X = (cen[rng.integers(k, size=N)] + rng.normal(0, 1, (N, D))).astype(np.float32)
        init = X[:k].copy()
        km = KMeans(n_clusters=k, init=init, n_init=1).fit(X)
        cases[name + "_shape"] = np.array([N, D, k, seed])
        cases[name + "_labels"] = km.labels_.astype(np.int32)
        cases[name + "_centers"] = km.cluster_centers_
        cases[name + "_inertia"] = np.float64(km.inertia_)
        cases[name + "_niter"] = np.int64(km.n_iter_)
    np.savez_compressed(f"{HERE}/kmeans_sklearn_dense.npz", **cases)
    print("dense kmeans goldens ok")




def make_seeded_kmeans_golden():
    """KMeans(n_clusters=k, random_state=int) -- k-means++ seeding with sklearn's RNG call sequence
    followed by Lloyd -- on uint8 rows (a blob set and one 1080p grid cell's worth of pixels, k = 8 as `-c 8`)."""
    from sklearn.cluster import KMeans, kmeans_plusplus
    cases = {}
    rng = np.random.default_rng(0)
    cen = rng.uniform(20, 230, (5, 4))
    blobs = np.clip(np.rint(cen[rng.integers(5, size=3000)] + rng.normal(0, 10, (3000, 4))), 0, 255).astype(np.uint8)
    cell = rng.integers(0, 256, (76 * 77, 4), dtype=np.uint8)
    for name, (X, k, rs) in {"blobs_k5_rs0": (blobs, 5, 0), "blobs_k5_rs42": (blobs, 5, 42), "cell_k8_rs3": (cell, 8, 3)}.items():
        km = KMeans(n_clusters=k, random_state=rs).fit(X)
        _, idx = kmeans_plusplus(X.astype(np.float64), k, random_state=rs)
        cases[name + "_X"] = X
        cases[name + "_k_rs"] = np.array([k, rs])
        cases[name + "_seed_idx"] = idx.astype(np.int64)
        cases[name + "_labels"] = km.labels_.astype(np.int32)
        cases[name + "_centers"] = km.cluster_centers_
        cases[name + "_inertia"] = np.float64(km.inertia_)
        cases[name + "_niter"] = np.int64(km.n_iter_)
    np.savez_compressed(f"{HERE}/kmeans_sklearn_seeded.npz", **cases)
    print("seeded kmeans goldens ok")


SWEEP_CASES = {   # name: (N, D, k, seed, dtype) -- the ends of BASELINE.json configs[4]'s sweep (D = 2 .. 5000)
    "u8_d2_k8": (20000, 2, 8, 21, "u8"), "f32_d2_k8": (20000, 2, 8, 22, "f32"),
    "f32_d2048_k16": (3000, 2048, 16, 23, "f32"), "f32_d5000_k8": (2000, 5000, 8, 24, "f32"),
    "u8_d5000_k8": (1500, 5000, 8, 25, "u8"), "f32_d2050_k12": (2500, 2050, 12, 26, "f32"),
}


def sweep_case(name):
    """regenerate the rows of a sweep case from its seed (the fixture stores only the sklearn results)"""
    N, D, k, seed, kind = SWEEP_CASES[name]
    rng = np.random.default_rng(seed)
    if kind == "u8":
        cen = rng.uniform(20, 230, (k, D))
        X = np.clip(np.rint(cen[rng.integers(k, size=N)] + rng.normal(0, 12, (N, D))), 0, 255).astype(np.uint8)
        init = X[:k].astype(np.float64)
    else:
        cen = rng.uniform(0, 8, (k, D))
        X = (cen[rng.integers(k, size=N)] + rng.normal(0, 1, (N, D))).astype(np.float32)
        init = X[:k].copy()
    return X, init


def make_sweep_kmeans_golden():
    """sklearn 1.9.0 KMeans(init=first k rows, n_init=1) at D = 2, 2048, 2050 (not a multiple of 4), 5000"""
    from sklearn.cluster import KMeans
    cases = {}
    for name in SWEEP_CASES:
        X, init = sweep_case(name)
        km = KMeans(n_clusters=init.shape[0], init=init, n_init=1).fit(X)
        cases[name + "_labels"] = km.labels_.astype(np.int32)
        cases[name + "_center_sums"] = km.cluster_centers_.astype(np.float64).sum(axis=1)     # [k]: keeps the fixture small
        cases[name + "_inertia"] = np.float64(km.inertia_)
        cases[name + "_niter"] = np.int64(km.n_iter_)
    np.savez_compressed(f"{HERE}/kmeans_sklearn_sweep.npz", **cases)
    print("sweep kmeans goldens ok")


MINIBATCH_CASES = {"lab_k8_rs0": (30000, 8, 0, 31), "lab_k4_rs7": (9000, 4, 7, 32), "lab_k16_rs3": (50000, 16, 3, 33)}


def minibatch_case(name):
    """LAB-like uint8 pixels: a few colour blobs on a smooth background (regenerated from the seed)"""
    n, k, rs, seed = MINIBATCH_CASES[name]
    rng = np.random.default_rng(seed)
    cen = rng.uniform(30, 220, (k + 2, 3))
    X = np.clip(np.rint(cen[rng.integers(k + 2, size=n)] + rng.normal(0, 9, (n, 3))), 0, 255).astype(np.uint8)
    return X, k, rs


def make_minibatch_golden():
    """SURVEY section 8f-4 (color-quantization/quant.py:18-19): sklearn 1.9.0 MiniBatchKMeans(n_clusters=k,
    random_state=rs).fit on uint8 pixels -- centres, labels, number of mini-batch steps, inertia"""
    from sklearn.cluster import MiniBatchKMeans
    cases = {}
    for name in MINIBATCH_CASES:
        X, k, rs = minibatch_case(name)
        clt = MiniBatchKMeans(n_clusters=k, random_state=rs).fit(X)
        cases[name + "_labels"] = clt.labels_.astype(np.int32)
        cases[name + "_centers"] = clt.cluster_centers_
        cases[name + "_inertia"] = np.float64(clt.inertia_)
        cases[name + "_nsteps"] = np.int64(clt.n_steps_)
    np.savez_compressed(f"{HERE}/minibatch_sklearn.npz", **cases)
    print("minibatch goldens ok")


if __name__ == "__main__":
    main()
