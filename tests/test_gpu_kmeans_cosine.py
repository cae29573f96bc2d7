"""GPU parity tests for the k-means / cosine stage and the reference-named drop-ins: the CUDA
path through the C-ABI against the oracle, the sklearn goldens and the reference's G1/G2/G4/
G5/G6 fixtures.  Bars (north_star): labels bit-exact given identical initial centres, inertia
within 1e-5 relative; cosine index sets bit-exact; CSV text identical for k = 1."""
import csv
import os

import numpy as np
import pytest
import torch

from oracle import grid_np as G
from oracle import kmeans_np as K
from oracle import viz_np as V
from tests.conftest import GOLDEN, unpack_images

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def km():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from opticalflowclustering_b200 import _lib, kmeans
    _lib.lib()
    return kmeans


def _hue_col(path):
    with open(path, encoding="utf-8-sig") as f:
        return np.array([int(r[1]) for r in csv.reader(f) if r])


@pytest.mark.parametrize("name", ["u8_d4_k3", "u8_d4_k8", "f32_d32_k16"])
def test_kmeans_fit_vs_sklearn_golden(km, name):
    z = np.load(os.path.join(GOLDEN, "kmeans_sklearn.npz"))
    X, init = z[name + "_X"], z[name + "_init"]
    labels, centres, inertia, n_iter = km.kmeans_fit(X, init)
    if name.startswith("u8"):
        assert (labels == z[name + "_labels"]).all()                     # bit-exact labels
        assert n_iter == int(z[name + "_niter"])
        assert np.abs(centres - z[name + "_centers"]).max() < 1e-9
        assert (km.predict(X, centres).cpu().numpy() == z[name + "_predict"]).all()
    else:
        # float32 data is worked in float32 (sklearn's sgemm chunks): the achieved count is printed and must be zero
        bad = int((labels != z[name + "_labels"]).sum())
        print(f"\n{name}: {bad} of {labels.size} labels differ from sklearn, n_iter {n_iter} vs {int(z[name + '_niter'])}")
        assert bad == 0 and n_iter == int(z[name + "_niter"])
        assert np.abs(centres - z[name + "_centers"]).max() < 1e-4
    assert abs(inertia - float(z[name + "_inertia"])) <= 1e-5 * float(z[name + "_inertia"])


@pytest.mark.parametrize("name", ["blobs_k5_rs0", "blobs_k5_rs42", "cell_k8_rs3"])
def test_kmeans_random_state_reproduces_sklearn(km, name):
    """KMeans(n_clusters=k, random_state=int) end to end against sklearn 1.9.0 (the reference's
    call with a seed): same k-means++ seed indices, labels and n_iter bit-exact"""
    z = np.load(os.path.join(GOLDEN, "kmeans_sklearn_seeded.npz"))
    X = z[name + "_X"]
    k, rs = (int(v) for v in z[name + "_k_rs"])
    _, idx = km.kmeans_plusplus(X, k, random_state=rs)
    assert (idx == z[name + "_seed_idx"]).all()
    clt = km.KMeans(n_clusters=k, random_state=rs).fit(X)
    assert (clt.labels_ == z[name + "_labels"]).all() and clt.n_iter_ == int(z[name + "_niter"])
    assert np.abs(clt.cluster_centers_ - z[name + "_centers"]).max() < 1e-9
    assert abs(clt.inertia_ - float(z[name + "_inertia"])) <= 1e-9 * float(z[name + "_inertia"])


@pytest.mark.parametrize("shape", [(200000, 4, 8, 1), (30001, 7, 5, 3), (20000, 32, 4, 2), (5852, 4, 8, 350)])
def test_fused_uint8_step_equals_two_kernel_iteration(km, shape, monkeypatch):
    """ofc_kmeans_step (E-step + exact integer sums in one pass) against ofc_kmeans_assign + ofc_kmeans_sums:
    labels, centres, inertia and n_iter bit-identical (the last shape is one 1080p frame's 350 cells)"""
    n, d, k, B = shape
    X = torch.from_numpy(np.random.default_rng(1).integers(0, 256, (B, n, d), dtype=np.uint8)).cuda()
    init = X[:, :k].double()
    monkeypatch.setenv("OFC_KMEANS_FUSED", "0")
    a = km.lloyd(X, init, max_iter=12)
    monkeypatch.setenv("OFC_KMEANS_FUSED", "1")
    b = km.lloyd(X, init, max_iter=12)
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_kmeans_large_n_properties(km):
    """1M x 4 uint8 rows, k = 8: too slow for the numpy oracle in full, so check (a) labels equal
    an independent fp64 torch restatement of the E-step on the final centres, (b) centres are the
    exact member means, (c) inertia equals the direct sum, (d) two runs are bit-identical."""
    g = torch.Generator().manual_seed(3)
    cen = torch.randint(20, 230, (8, 4), generator=g).double()
    X = (cen[torch.randint(0, 8, (1_000_000,), generator=g)] + 12 * torch.randn(1_000_000, 4, generator=g, dtype=torch.float64))
    X = X.round().clamp(0, 255).to(torch.uint8).cuda()
    init = X[:8].double()
    l1, c1, i1, n1 = km.lloyd(X, init, tol=0.0)          # tol = 0: stop only when the labels repeat
    l2, c2, i2, n2 = km.lloyd(X, init, tol=0.0)
    assert int(n1) < 300
    assert torch.equal(l1, l2) and torch.equal(c1, c2) and float(i1) == float(i2) and int(n1) == int(n2)
    Xd = X.double()
    mean = Xd.mean(0)
    d = ((c1 - mean) ** 2).sum(1)[None, :] - 2.0 * (Xd - mean) @ (c1 - mean).T
    ref = d.argmin(1).to(torch.int32)
    assert (ref != l1).sum().item() <= 2                                   # near-ties only
    for j in range(8):
        m = Xd[l1 == j].mean(0)
        assert (m - c1[j]).abs().max().item() < 1e-9
    direct = ((Xd - c1[l1.long()]) ** 2).sum().item()
    assert abs(direct - float(i1)) <= 1e-9 * direct


def test_kmeans_batched_cells_k3_vs_oracle(km):
    z = np.load(os.path.join(GOLDEN, "g23_cells.npz"))
    cells = z["cells"][0, 100:130]
    X = np.stack([G.preprocess_image(c.copy()).reshape(-1, 4) for c in cells])
    init = np.stack([np.unique(x, axis=0)[[0, len(np.unique(x, axis=0)) // 2, -1]] for x in X]).astype(np.float64)
    labels, centres, inertia, n_iter = km.lloyd(X, init)
    for b in range(len(X)):
        w = K.kmeans_fit(X[b], init[b])
        assert (labels[b].cpu().numpy() == w[0]).all()
        assert int(n_iter[b]) == w[3]
        assert np.abs(centres[b].cpu().numpy() - w[1]).max() < 1e-9
        assert abs(float(inertia[b]) - w[2]) <= 1e-9 * max(w[2], 1.0)


def test_kmeans_errors(km):
    with pytest.raises(ValueError):
        km.KMeans(n_clusters=4).fit(np.zeros((3, 4), np.uint8))
    with pytest.raises(ValueError):
        km.lloyd(np.zeros((10, 4), np.uint8), np.zeros((2, 5)))


def test_color_kmeans_g1_csv_text(km, tmp_path, monkeypatch):
    """G1: color_kmeans.py end to end, k = 1 -> cluster_centers.csv text identical (CRLF rows)."""
    from opticalflowclustering_b200 import color_kmeans as ck
    z = np.load(os.path.join(GOLDEN, "g1_images.npz"))
    monkeypatch.chdir(tmp_path)
    open("cluster_centers.csv", "w").close()
    for im, name in zip(unpack_images(z), z["names"]):
        rgb = np.ascontiguousarray(im[..., ::-1])            # read_image's BGR -> RGB
        ck.cluster_colors(ck.preprocess_image(rgb), 1, "some/dir/" + str(name), "cluster_centers.csv")
    got = open("cluster_centers.csv", newline="").read()
    want = open(os.path.join(GOLDEN, "g1_cluster_centers.csv"), newline="").read()
    assert got.replace("\r\n", "\n") == want.replace("\r\n", "\n")
    assert "\r\n" in got                                       # csv.writer line ends, like the reference


def test_preprocess_image_in_place(km):
    from opticalflowclustering_b200 import color_kmeans as ck
    rng = np.random.default_rng(0)
    im = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    want_in = im.copy()
    want = G.preprocess_image(want_in)
    got = ck.preprocess_image(im)
    assert (got == want).all() and (im == want_in).all()


def test_g4_hues_through_dropin(km):
    from opticalflowclustering_b200 import KmeanGrids as kg
    z = np.load(os.path.join(GOLDEN, "g4_images.npz"))
    for im, want in list(zip(unpack_images(z), z["hues"]))[:12]:
        rgb = np.ascontiguousarray(im[..., ::-1])
        c, hue = kg.cluster_colors(kg.preprocess_image(rgb), 1, "x", os.devnull)
        assert hue == want


def test_kmeangrids_frame_hues_and_outcsv(km, tmp_path):
    """overlayGridAndComputeAvgColor + the main loop's per-cell k = 1 hues on a synthetic
    visualisation frame vs the oracle's restatement of the reference loop; OutCSV row text."""
    from opticalflowclustering_b200 import KmeanGrids as kg
    rng = np.random.default_rng(5)
    frame = rng.integers(0, 256, (714, 1275, 3), dtype=np.uint8)
    ref = frame.copy()
    _, _, rois = G.grid_mean_hues(ref, 14, 25)
    ref_lines = ref.copy()                                             # preprocess_image mutates the ROI views
    want = [G.cluster_colors_k1(G.preprocess_image(r))[1] for r in rois]
    kg.image_dict.clear()
    kg.overlayGridAndComputeAvgColor(2, frame, kg.GRID_PARAMS, "unused.csv", "clip.mp4")
    assert len(kg.image_dict) == 350 and kg.image_dict["2/1"].base is not None
    assert (frame == ref_lines).all()                                  # grid lines as the reference leaves them
    hues = kg.frame_hues("2", [str(i) for i in range(1, 351)], 1)
    assert hues == [int(h) for h in want]
    # k = 1 through the generic path (image_dict ROI -> preprocess -> Lloyd kernels) agrees too
    kg.frame_results.clear()
    assert kg.frame_hues("2", ["1", "26", "350"], 1) == [int(want[0]), int(want[25]), int(want[349])]
    p = tmp_path / "out.csv"
    kg.write_outcsv_row(str(p), hues, True)
    kg.write_outcsv_row(str(p), hues, False)
    rows = list(csv.reader(open(p)))
    assert rows[0][0] == "cell_0" and rows[0][-1] == "cell_349" and len(rows) == 3 and rows[1] == [str(h) for h in hues]


def test_lloyd_cells_one_launch_vs_oracle(km):
    """per-cell device-resident Lloyd runs (one launch for all cells) == oracle with the same initial centres"""
    z = np.load(os.path.join(GOLDEN, "g23_cells.npz"))
    cells = z["cells"][:2].reshape(-1, 51, 51, 3)
    X = np.stack([G.preprocess_image(c.copy()).reshape(-1, 4) for c in cells])
    # cells with fewer distinct rows than clusters are degenerate: which point an empty cluster takes is then
    # decided by rounding noise of sklearn's centred sums (all distances are "zero"), not by the algorithm
    X = X[[i for i, x in enumerate(X) if len(np.unique(x, axis=0)) >= 8]][:48]
    assert len(X) >= 20
    init = np.stack([np.unique(x, axis=0)[[0, len(np.unique(x, axis=0)) // 3, 2 * len(np.unique(x, axis=0)) // 3, -1]] for x in X]).astype(np.float64)
    lab, cen, inertia, n_iter, counts = km.lloyd_cells(X, 4, init=init)
    for b in range(len(X)):
        w = K.kmeans_fit(X[b], init[b])
        assert (lab[b].cpu().numpy() == w[0]).all() and int(n_iter[b]) == w[3]
        assert np.abs(cen[b].cpu().numpy() - w[1]).max() < 1e-9
        assert abs(float(inertia[b]) - w[2]) <= 1e-9 * max(w[2], 1.0)
        assert (counts[b].cpu().numpy() == np.bincount(w[0], minlength=4)).all()


def test_kmeangrids_k2_batched_dominant_hue(km):
    """k = 2 through KmeanGrids.frame_hues: cells painted with a 70/30 mix of two flat colours -> the dominant
    colour's hue, whatever the k-means++ draw"""
    from opticalflowclustering_b200 import KmeanGrids as kg
    rng = np.random.default_rng(8)
    kg.image_dict.clear()
    kg.frame_results.clear()
    want = []
    for c in range(1, 21):
        major = rng.integers(60, 255, 3)
        minor = (major.astype(int) + 120) % 200 + 40
        roi = np.empty((40, 40, 3), np.uint8)
        roi[:] = major
        roi[:12] = minor                                                   # 30 % of the pixels
        kg.image_dict[f"5/{c}"] = roi
        want.append(int(V.bgr2hsv_u8(major.astype(np.uint8)[None, None])[0, 0, 0]))
    assert kg.frame_hues("5", [str(c) for c in range(1, 21)], 2) == want


def test_drawgrids_csv_row(km, tmp_path):
    from opticalflowclustering_b200 import drawGridsAndOutputCSV as dg
    rng = np.random.default_rng(6)
    frame = rng.integers(0, 256, (200, 300, 3), dtype=np.uint8)
    ref = frame.copy()
    _, hues, _ = G.grid_mean_hues(ref, 10, 10)
    p = tmp_path / "rgb_values.csv"
    dg.overlayGridAndComputeAvgColor(2, frame, dg.GRID_PARAMS, str(p))
    dg.overlayGridAndComputeAvgColor(3, frame.copy(), dg.GRID_PARAMS, str(p))
    rows = list(csv.reader(open(p)))
    assert len(rows) == 3 and rows[0][:2] == ["cell_0", "cell_1"]
    assert rows[1] == [str(float(h)) for h in hues]
    assert (frame == ref).all()                               # same white grid lines as the reference


def test_cosine_goldens_g5_g6(km, capsys, tmp_path, monkeypatch):
    from opticalflowclustering_b200 import computeVectorDistance as cvd
    from opticalflowclustering_b200 import findCosineDifferentVectors as fc
    best, frame = fc.main([os.path.join(GOLDEN, "bounce.csv"), os.path.join(GOLDEN, "601_3_3_cropped.csv")])
    out = capsys.readouterr().out.splitlines()
    assert out[0] == "Vector sizes are:  16 75"
    assert out[1] == "Maximum cosine similarity: 0.91448231723348"
    assert out[2] == "Minimum sum of squared differences: 0" and out[3] == "Max frame: 24"
    short = _hue_col(os.path.join(GOLDEN, "bounce.csv"))
    long_ = _hue_col(os.path.join(GOLDEN, "cropped_trimmed2.csv"))
    from opticalflowclustering_b200 import cosine
    b2, f2, sims = cosine.sliding_cosine(short, long_, return_sims=True)
    ob, of = G.sliding_cosine(short, long_)
    assert (b2, f2) == (ob, of) and f2 == 7
    ref = np.array([G.cosine_similarity(short, long_[i:i + 16]) for i in range(len(sims))])
    assert (sims == ref).all()                                                    # bit-exact
    sim, row, dist = cvd.main(os.path.join(GOLDEN, "file1.csv"), os.path.join(GOLDEN, "file2.csv"))
    out = capsys.readouterr().out
    assert out.startswith("[[1.]]\n") and "Euclidean distance: 0.0" in out
    _, orow, _ = G.vector_distance(_hue_col(os.path.join(GOLDEN, "file1.csv")), _hue_col(os.path.join(GOLDEN, "file2.csv")))
    assert np.array_equal(row, orow, equal_nan=True)


def test_row_cosine_large(km):
    from opticalflowclustering_b200 import cosine
    g = torch.Generator().manual_seed(1)
    X = torch.randint(0, 180, (1_000_000, 16), generator=g, dtype=torch.uint8).cuda()
    q = torch.randint(0, 180, (16,), generator=g).double()
    out = cosine.row_cosine(X, q)
    Xd = X.double()
    ref = (Xd @ q.cuda()) / (Xd.norm(dim=1) * q.norm())
    assert (out - ref).abs().max().item() < 1e-14
    sub = X[:64].cpu().numpy().astype(np.float64)
    want = np.array([G.cosine_similarity(r, q.numpy()) for r in sub])
    assert (out[:64].cpu().numpy() == want).all()                                 # integer data: exact


def test_sliding_cosine_large_properties(km):
    from opticalflowclustering_b200 import cosine
    rng = np.random.default_rng(2)
    long_ = rng.integers(0, 180, 1_000_000)
    short = long_[777_000:777_016].copy()
    long_[123_000:123_016] = short                                                # two exact matches: last wins
    best, frame, sims = cosine.sliding_cosine(short, long_, return_sims=True)
    assert frame == 777_000 and abs(best - 1.0) < 1e-15
    assert sims[123_000] == sims[777_000] == sims.max()


# ---- per-cell KMeans(k > 1): the shared-memory / filtered kernel and the device-resident pipeline --------------
def _cells_both(km, X, k, monkeypatch, **kw):
    monkeypatch.setenv("OFC_CELLS_FAST", "1")
    fast = km.lloyd_cells(X, k, **kw)
    monkeypatch.setenv("OFC_CELLS_FAST", "0")
    slow = km.lloyd_cells(X, k, **kw)
    monkeypatch.delenv("OFC_CELLS_FAST")
    return fast, slow


@pytest.mark.parametrize("k", [2, 8, 16])
def test_cells_fast_kernel_bit_identical_to_first_kernel_full_frame(km, k, monkeypatch):
    """one 1080p frame's worth of cells (350 x 5852 px, the BASELINE shape): float32-filtered E-step with float64
    re-evaluation + incremental integer M-step == the all-float64 kernel, every output, bit for bit"""
    g = torch.Generator().manual_seed(k)
    X = torch.randint(0, 256, (350, 76 * 77, 4), generator=g).to(torch.uint8)
    X[:40, :, 3] = 255                                               # some cells on a 3-D slice
    X[40:60] = (X[40:60] // 64) * 64                                  # heavy duplicates: exact ties
    X = X.cuda()
    for kw in ({"seed": 11}, {"init": X[:, :k].double()}):
        fast, slow = _cells_both(km, X, k, monkeypatch, **kw)
        for name, u, v in zip(("labels", "centres", "inertia", "n_iter", "counts"), fast, slow):
            assert torch.equal(u, v), name


def test_cells_fast_kernel_vs_numpy_oracle_with_seeding(km):
    """device-resident seeding + fit + dominant hue against oracle/kmeans_np.py (cells_fit, dominant_centre_hue)"""
    from opticalflowclustering_b200 import grid
    rng = np.random.default_rng(71)
    H, W, rows, cols, k = 96, 120, 4, 5, 3
    frames = np.zeros((2, H, W, 3), np.uint8)
    for f in range(2):
        for cy in range(rows):
            for cx in range(cols):
                cen = rng.uniform(0, 255, (k, 3))
                cell = cen[rng.integers(k, size=(H // rows, W // cols))] + rng.normal(0, 12, (H // rows, W // cols, 3))
                frames[f, cy * 24:(cy + 1) * 24, cx * 24:(cx + 1) * 24] = np.clip(np.rint(cell), 0, 255)
    out = grid.grid_kmeans_cells(torch.from_numpy(frames).cuda(), k, rows, cols, seed=5, first_frame=7, want_centres=True)
    for f in range(2):
        fr = frames[f].copy()
        _, _, rois = G.grid_mean_hues(fr, rows, cols)                       # ROI views with the white lines drawn
        for c, roi in enumerate(rois):
            X = G.preprocess_image(roi.copy()).reshape(-1, 4)
            lab, cen, inertia, n_iter, idx = K.cells_fit(X, k, 5, (7 + f) * rows * cols + c)
            assert int(out["n_iter"][f, c]) == n_iter
            assert np.abs(out["centres"][f, c].cpu().numpy() - cen).max() <= 1e-9
            assert (out["counts"][f, c].cpu().numpy() == np.bincount(lab, minlength=k)).all()
            want_c, want_h = K.dominant_centre_hue(X, lab, cen)
            assert (out["dom_centre"][f, c].cpu().numpy() == want_c.astype(np.uint8)).all()
            assert int(out["dom_hue"][f, c]) == want_h


def test_clip_pipeline_k8_device_resident(km):
    """ClipPipeline(n_clusters=8): flow -> visualisation -> per-cell KMeans(8) hues without leaving the device;
    equals the oracle run on the pipeline's own visualisation, and does not depend on the chunking"""
    from opticalflowclustering_b200.pipeline import ClipPipeline
    from opticalflowclustering_b200.synthetic import synthetic_clip
    H, W, T, rows, cols, k = 144, 200, 5, 6, 8, 8
    clip = synthetic_clip(T, H, W, seed=13)
    pipe = ClipPipeline(W, H, chunk_frames=T, rows=rows, cols=cols, n_clusters=k, kmeans_seed=3)
    pipe.run_chunk(clip.cuda())
    viz = pipe.viz.cpu().numpy()
    hue = pipe.km_hue.cpu().numpy()
    nit = pipe.km_n_iter.cpu().numpy()
    for p in (0, T - 2):
        fr = viz[p].copy()
        _, _, rois = G.grid_mean_hues(fr, rows, cols)
        for c in range(0, rows * cols, 5):
            X = G.preprocess_image(rois[c].copy()).reshape(-1, 4)
            lab, cen, inertia, n_iter, idx = K.cells_fit(X, k, 3, p * rows * cols + c)
            assert int(nit[p, c]) == n_iter
            assert int(hue[p, c]) == K.dominant_centre_hue(X, lab, cen)[1]
    # two chunks (3 + 2 new frames) give the same rows as one
    pipe2 = ClipPipeline(W, H, chunk_frames=3, rows=rows, cols=cols, n_clusters=k, kmeans_seed=3)
    res = pipe2.process_clip(clip)
    assert (res["km_hue"].numpy() == hue).all()
    # k = 1 keeps riding the grid pass
    pipe1 = ClipPipeline(W, H, chunk_frames=T, rows=rows, cols=cols)
    pipe1.run_chunk(clip.cuda())
    assert (pipe1.avg_hue.cpu().numpy() == pipe.avg_hue.cpu().numpy()).all()


# ---- the ends of BASELINE configs[4]'s sweep: D = 2, 2048, 2050 (not a multiple of 4), 5000; k > 4096 ------------------
@pytest.mark.parametrize("name", ["u8_d2_k8", "f32_d2_k8", "f32_d2048_k16", "f32_d5000_k8", "u8_d5000_k8", "f32_d2050_k12"])
def test_kmeans_sweep_ends_vs_sklearn_golden(km, name):
    """labels bit-exact (the mismatch count is printed), same n_iter, inertia within 1e-5 relative, centre row sums"""
    import sys
    sys.path.insert(0, GOLDEN)
    import make_golden as MG
    z = np.load(os.path.join(GOLDEN, "kmeans_sklearn_sweep.npz"))
    X, init = MG.sweep_case(name)
    labels, centres, inertia, n_iter = km.kmeans_fit(X, init)
    bad = int((labels != z[name + "_labels"]).sum())
    print(f"\n{name}: {bad} of {labels.size} labels differ from sklearn, n_iter {n_iter} vs {int(z[name + '_niter'])}")
    assert bad == 0 and n_iter == int(z[name + "_niter"])
    assert abs(inertia - float(z[name + "_inertia"])) <= 1e-5 * float(z[name + "_inertia"])
    want = z[name + "_center_sums"]
    assert np.abs(centres.sum(axis=1) - want).max() <= 1e-4 * max(1.0, np.abs(want).max())


def test_kmeans_more_than_4096_clusters(km):
    """k = 6000 (past the tensor-core path's and the first relocation kernel's limits): the generic kernels take over;
    labels equal the oracle's for the same initial centres.  float64 rows: the generic E-step sums lane-strided partial
    products, so rows EXACTLY equidistant from two centres (integer lattice data) may resolve differently from the
    sequential BLAS order there; continuous data has no such rows"""
    rng = np.random.default_rng(9)
    N, D, k = 24000, 8, 6000
    X = rng.normal(0, 50, (N, D))
    init = X[:k].copy()
    labels, centres, inertia, n_iter = km.kmeans_fit(X, init, max_iter=2)
    w = K.kmeans_fit(X, init, max_iter=2)
    assert (labels == w[0]).all() and n_iter == w[3]
    assert abs(inertia - w[2]) <= 1e-9 * w[2]
