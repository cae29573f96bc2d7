"""oracle/reference_chain.py (the CPU baseline) reproduces the reference's own outputs."""
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN, have_cv2

pytestmark = pytest.mark.skipif(not have_cv2(), reason="cv2 not importable")


def test_flow_state_matches_reference_compute():
    from oracle import reference_chain as RC
    z = np.load(os.path.join(GOLDEN, "flow_135x240.npz"))
    st = RC.FlowState(z["clip"][0].copy())
    for p in range(2):
        assert (st.compute(z["clip"][p + 1].copy()) == z["viz"][p]).all()


def test_chain_matches_oracle_restatement():
    pytest.importorskip("sklearn")
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import grid_np as G
    from oracle import reference_chain as RC
    z = np.load(os.path.join(GOLDEN, "flow_135x240.npz"))
    avg, km = RC.run_frames(z["clip"].copy(), 1, 6, 8)
    for p in range(2):
        fr = z["viz"][p].copy()
        _, hue, rois = G.grid_mean_hues(fr, 6, 8)
        assert (avg[p] == hue).all()
        kh = [G.cluster_colors_k1(G.preprocess_image(r))[1] for r in rois]
        assert (km[p] == np.array(kh)).all()
