"""Pin the numpy Farneback / visualisation oracle against cv2 (live and golden)."""
import os

import numpy as np
import pytest

from oracle import farneback_np as FB
from oracle import viz_np as V
from tests.conftest import GOLDEN, have_cv2


@pytest.mark.parametrize("name", ["flow_96x128", "flow_135x240"])
def test_oracle_flow_vs_golden_cv2(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    for p in range(2):
        mine = FB.calc_optical_flow_farneback(z["gray"][p], z["gray"][p + 1])
        epe = np.linalg.norm(mine - z["flow"][p], axis=-1)
        assert epe.mean() < 2e-6 and epe.max() < 1e-4, (epe.mean(), epe.max())


@pytest.mark.skipif(not have_cv2(), reason="cv2 not importable")
@pytest.mark.parametrize("kw", [dict(), dict(levels=1, winsize=9, iterations=2),
                                dict(pyr_scale=0.6, levels=2, winsize=11, poly_n=7, poly_sigma=1.5)])
def test_oracle_flow_vs_live_cv2(kw):
    import cv2
    from opticalflowclustering_b200.synthetic import synthetic_clip
    clip = synthetic_clip(2, 101, 150, seed=9).numpy()
    g = np.stack([V.bgr2gray(f) for f in clip])
    a = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2)
    a.update(kw)
    ref = cv2.calcOpticalFlowFarneback(g[0], g[1], None, a["pyr_scale"], a["levels"], a["winsize"], a["iterations"],
                                       a["poly_n"], a["poly_sigma"], 0)
    mine = FB.calc_optical_flow_farneback(g[0], g[1], **a)
    epe = np.linalg.norm(mine - ref, axis=-1)
    assert epe.mean() < 2e-6 and epe.max() < 2e-4, (epe.mean(), epe.max())


@pytest.mark.skipif(not have_cv2(), reason="cv2 not importable")
@pytest.mark.parametrize("flags,kw", [(256, dict()), (4, dict()), (260, dict(winsize=9)),
                                      (4, dict(pyr_scale=0.6, levels=2, winsize=11)), (256, dict(winsize=21, levels=1))])
def test_oracle_flow_flags_vs_live_cv2(flags, kw):
    """OPTFLOW_FARNEBACK_GAUSSIAN (256) and OPTFLOW_USE_INITIAL_FLOW (4)"""
    import cv2
    from opticalflowclustering_b200.synthetic import synthetic_clip
    clip = synthetic_clip(3, 101, 150, seed=9).numpy()
    g = np.stack([V.bgr2gray(f) for f in clip])
    a = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2)
    a.update(kw)
    init = cv2.calcOpticalFlowFarneback(g[0], g[1], None, 0.5, 3, 15, 3, 5, 1.2, 0)
    ref = cv2.calcOpticalFlowFarneback(g[1], g[2], init.copy(), a["pyr_scale"], a["levels"], a["winsize"],
                                       a["iterations"], a["poly_n"], a["poly_sigma"], flags)
    mine = FB.calc_optical_flow_farneback(g[1], g[2], init.copy(), flags=flags, **a)
    epe = np.linalg.norm(mine - ref, axis=-1)
    assert epe.mean() < 2e-6 and epe.max() < 2e-4, (epe.mean(), epe.max())


@pytest.mark.skipif(not have_cv2(), reason="cv2 not importable")
@pytest.mark.parametrize("shape", [(96, 128, 48, 64), (135, 240, 34, 60), (101, 150, 25, 38), (77, 131, 39, 66)])
def test_oracle_resize_area_vs_live_cv2(shape):
    import cv2
    h, w, dh, dw = shape
    f = (np.random.default_rng(1).standard_normal((h, w, 2)) * 3).astype(np.float32)
    assert (FB.resize_area_f32(f, dw, dh) == cv2.resize(f, (dw, dh), interpolation=cv2.INTER_AREA)).all()


def test_pyramid_plan_levels_plus_one():
    # Q1: levels=3 means 4 scales; cvRound sizes; ksize 19/9/3/3 at 1080p
    plan = FB.pyramid_plan(1920, 1080, 0.5, 3)
    assert [(p[4], p[5], p[3]) for p in plan] == [(240, 135, 19), (480, 270, 9), (960, 540, 3), (1920, 1080, 3)]
    plan4k = FB.pyramid_plan(3840, 2160, 0.5, 5)
    assert (plan4k[0][4], plan4k[0][5], plan4k[0][3]) == (120, 68, 79)
    # small image: levels cropped while the coarse side stays >= 32
    assert len(FB.pyramid_plan(128, 96, 0.5, 3)) == 2


@pytest.mark.parametrize("name", ["flow_96x128", "flow_270x480"])
def test_oracle_viz_chain_vs_reference_golden(name):
    """golden viz = the reference's own ComputeOpticalFLow.compute output"""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    assert (np.stack([V.bgr2gray(f) for f in z["clip"]]) == z["gray"]).all()
    for p in range(2):
        bgr, _ = V.flow_to_bgr(z["flow"][p])
        assert (bgr == z["viz"][p]).all()


@pytest.mark.skipif(not have_cv2(), reason="cv2 not importable")
def test_oracle_colour_formulas_vs_live_cv2():
    import cv2
    rng = np.random.default_rng(0)
    bgr = rng.integers(0, 256, (600, 700, 3), dtype=np.uint8)
    assert (V.bgr2gray(bgr) == cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)).all()
    assert (V.bgr2hsv_u8(bgr) == cv2.cvtColor(bgr, cv2.COLOR_BGR2HSV)).all()
    x = (rng.standard_normal((400, 500)) * rng.choice([0.01, 1, 30], (400, 500))).astype(np.float32)
    y = (rng.standard_normal((400, 500)) * rng.choice([0.01, 1, 30], (400, 500))).astype(np.float32)
    x[0, :10] = 0
    y[0, :5] = 0
    m, a = cv2.cartToPolar(x, y)
    mm, aa = V.cart_to_polar(x, y)
    assert (m == mm).all() and (a == aa).all()
    ref = np.zeros(x.shape, np.uint8)
    ref[...] = cv2.normalize(m, None, 0, 255, cv2.NORM_MINMAX)
    assert (V.normalize_minmax_u8(m) == ref).all()
    H, Vv = np.meshgrid(np.arange(0, 181), np.arange(256), indexing="ij")
    hsv = np.stack([H, np.full_like(H, 255), Vv], -1).astype(np.uint8)
    assert (V.hsv2bgr_s255(hsv[..., 0], hsv[..., 2]) == cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR)).all()
