"""End-to-end GPU tests of the reference-named CLIs on a real (synthetic, lossy-encoded) video file:
`KmeanGrids.py -d OutImgs/<v> -c 1 -f x.csv --noyolo --nocontour --path <v>.mp4` and
`computeOpticalFlow.py -i <v>.mp4`, compared with the oracle's restatement of the reference's own call
chain (oracle/reference_chain.py: cv2 + sklearn as the reference calls them) on the same decoded frames."""
import csv
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

cv2 = pytest.importorskip("cv2")


def _write_clip(path, n=6, H=280, W=400, seed=17):
    from opticalflowclustering_b200.synthetic import synthetic_clip
    clip = synthetic_clip(n, H, W, seed=seed).numpy()
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 25.0, (W, H))
    if not wr.isOpened():
        pytest.skip("cv2.VideoWriter cannot encode in this environment")
    for f in clip:
        wr.write(f)
    wr.release()
    cap = cv2.VideoCapture(path)
    frames = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        frames.append(f)
    cap.release()
    assert len(frames) == n
    return np.stack(frames)


def test_kmeangrids_cli_matches_reference_chain(tmp_path, monkeypatch):
    assert torch.cuda.is_available()
    from opticalflowclustering_b200 import KmeanGrids as kg
    from oracle import reference_chain as RC
    monkeypatch.chdir(tmp_path)
    frames = _write_clip("clipA.mp4")
    n_pairs = len(frames) - 1
    # the reference lists OutImgs/<video>/<frame>/*.png (written by an earlier drawGridsAndOutputCSVChange.py run)
    # only for the names: provide them
    for fr in range(2, 2 + n_pairs):
        os.makedirs(f"OutImgs/clipA/{fr}")
        for c in range(1, 351):
            open(f"OutImgs/clipA/{fr}/{c}.png", "wb").close()
    kg.image_dict.clear()
    kg.frame_results.clear()
    kg.main(["-d", "OutImgs/clipA", "-c", "1", "-f", "x.csv", "--noyolo", "--nocontour", "--path", "clipA.mp4"])
    rows = list(csv.reader(open("OutCSV/clipA.csv")))
    assert rows[0] == [f"cell_{i}" for i in range(350)] and len(rows) == 1 + n_pairs
    got = np.array([[int(v) for v in r] for r in rows[1:]])
    _, want = RC.run_frames(frames.copy(), n_clusters=1)
    # flow differs from cv2 by ~1e-7 px: a uint8 truncation boundary can flip a pixel and with it, rarely,
    # a cell's rounded mean -- allow a handful of 350*n cells to differ
    mism = (got != want)
    assert mism.mean() <= 0.01, mism.sum()
    assert len(kg.image_dict) == 350 * n_pairs


def test_compute_optical_flow_cli(tmp_path, monkeypatch):
    from opticalflowclustering_b200 import computeOpticalFlow as cof
    from oracle import viz_np as V
    monkeypatch.chdir(tmp_path)
    frames = _write_clip("clipB.mp4", n=7)
    means = cof.run("clipB.mp4", chunk_frames=4)
    assert len(means) == len(frames) - 1
    rows = list(csv.reader(open("clipB.mp4_opticalFlow.csv")))
    assert rows[0] == ["", "Frame", "Average Magnitude"] and len(rows) == len(frames)
    gray = [cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in frames]
    for t in range(len(frames) - 1):
        flow = cv2.calcOpticalFlowFarneback(gray[t], gray[t + 1], None, 0.5, 3, 15, 3, 5, 1.2, 0)
        mag, _ = cv2.cartToPolar(flow[..., 0], flow[..., 1])
        assert abs(float(rows[1 + t][2]) - float(np.mean(mag))) <= 1e-4 * max(float(np.mean(mag)), 1e-3)
    cap = cv2.VideoCapture("clipB.mp4onlyOpticalflow.mp4")
    assert cap.isOpened() and int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == len(frames) - 1


def test_drawgrids_change_then_color_kmeans_change(tmp_path, monkeypatch):
    """The two-step workflow that produced the reference's fixtures: drawGridsAndOutputCSVChange.py writes the
    cell PNGs + <video>_rgb_values.csv, color_kmeansChange.py -d OutImgs/<video>/ -c 1 clusters them."""
    from opticalflowclustering_b200 import color_kmeansChange as ckc
    from opticalflowclustering_b200 import drawGridsAndOutputCSVChange as dgc
    from oracle import grid_np as G
    from oracle import reference_chain as RC
    monkeypatch.chdir(tmp_path)
    frames = _write_clip("clipC.mp4", n=3)
    dgc.main(["--noyolo", "--nocontour", "--path", "clipC.mp4"])
    assert sorted(os.listdir("OutImgs/clipC"), key=int) == ["2", "3"] and len(os.listdir("OutImgs/clipC/2")) == 350
    rows = list(csv.reader(open("clipC.mp4_rgb_values.csv")))
    assert len(rows) == 3 and rows[0][0] == "cell_0"
    # reference chain on the same frames: grid hues of the flow visualisation
    st = RC.FlowState(frames[0])
    viz = st.compute(frames[1])
    hues, rois = RC.grid_pass(viz.copy())
    got = np.array([float(v) for v in rows[1]])
    assert (got != hues).mean() <= 0.01
    # a saved cell = ROI of the lined frame: white first row / column
    cell = cv2.imread("OutImgs/clipC/2/27.png")
    assert (cell[0] == 255).all() and (cell[:, 0] == 255).all()
    open("cluster_centers.csv", "w").close()
    ckc.main(["-d", "OutImgs/clipC/", "-c", "1", "-f", "cluster_centers.csv"])
    out = list(csv.reader(open("cluster_centers.csv", newline="")))
    assert out[0] == ["File name", "Cluster 1", "HSV Cluster 1", "Hue 0"] and len(out) == 1 + 700
    assert out[1][0] == "2/1.png" and out[350][0] == "2/350.png" and out[351][0] == "3/1.png"
    for k in (1, 27, 350):
        im = cv2.cvtColor(cv2.imread(f"OutImgs/clipC/2/{k}.png"), cv2.COLOR_BGR2RGB)
        c, h = G.cluster_colors_k1(G.preprocess_image(im))
        assert out[k][1] == str(c) and out[k][3] == str(h)
