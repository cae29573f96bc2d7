"""GPU parity tests of the tensor-core k-means path (csrc/tc_kmeans.cu: 3xTF32 tcgen05 E-step + CSR M-step),
called through the C-ABI.  Bars: labels bit-identical to the float32 CUDA-core E-step on the same centred
rows (the tensor cores only filter; near-ties are re-evaluated in float32), counts exact, float64 sums equal to
an index_add restatement; whole fits against scikit-learn 1.9.0 goldens (float32 data: >= 99.95 % labels,
inertia within 1e-5 relative -- the bar of the existing float32 case)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tc():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    spec = importlib.util.spec_from_file_location("tc_check", os.path.join(ROOT, "tools", "tc_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("shape", [(1000, 36, 2), (4096, 64, 64), (5000, 100, 40), (20000, 128, 300), (50000, 512, 1024),
                                   (100000, 64, 256)])
def test_tc_steps_equal_float32_kernels(tc, shape):
    """ragged n (not a multiple of 128), d not a multiple of 32, k not a multiple of the 64/128/256 centre tile"""
    assert tc.run(*shape)


def test_tc_steps_uniform_data_and_tied_centres(tc):
    """unclustered data (every distance close) and duplicated / 1e-6-perturbed centres: every row is a near-tie
    and must come out of the float32 re-evaluation with the lowest-index rule"""
    assert tc.run(20000, 64, 64, blobs=False)
    assert tc.run(30000, 96, 128, ties=True)
    assert tc.run(30000, 256, 512, ties=True, blobs=False)


@pytest.mark.parametrize("shape", [(3000, 36, 8), (20000, 352, 8), (50000, 128, 256), (30000, 352, 1024)])
def test_tc_steps_uint8_equal_float64_kernels(tc, shape):
    """uint8 rows: labels bit-identical to the float64 CUDA-core E-step (hence to sklearn given the centres),
    integer sums and counts exact"""
    assert tc.run_u8(*shape)


def test_tc_steps_uint8_tied_centres(tc):
    assert tc.run_u8(20000, 64, 128, ties=True)


def test_lloyd_uint8_dense_tensor_core_path_on_off(monkeypatch):
    """whole uint8 fits (hue vectors per frame: d = 352, k = 64) through the float64 kernels and the tensor-core
    path: everything bit-identical, because both E-steps give the same labels and the sums are exact integers"""
    from opticalflowclustering_b200 import kmeans
    g = torch.Generator().manual_seed(8)
    cen = torch.rand((64, 352), generator=g) * 150 + 20
    X = (cen[torch.randint(0, 64, (20000,), generator=g)] + 10 * torch.randn((20000, 352), generator=g))
    X = X.round().clamp(0, 255).to(torch.uint8).cuda()
    init = X[:64].double()
    monkeypatch.setenv("OFC_KMEANS_TC", "0")
    a = kmeans.lloyd(X, init)
    monkeypatch.setenv("OFC_KMEANS_TC", "1")
    b = kmeans.lloyd(X, init)
    assert all(torch.equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("name", ["f32_d64_k32", "f32_d96_k40"])
def test_lloyd_dense_float32_vs_sklearn_golden(name):
    from opticalflowclustering_b200 import kmeans
    z = np.load(os.path.join(GOLDEN, "kmeans_sklearn_dense.npz"))
    N, D, k, seed = (int(v) for v in z[name + "_shape"])
    rng = np.random.default_rng(seed)                       # same stream as tests/golden/make_golden.py
    cen = rng.uniform(0, 8, (k, D))
    X = (cen[rng.integers(k, size=N)] + rng.normal(0, 1, (N, D))).astype(np.float32)
    labels, centres, inertia, n_iter = kmeans.kmeans_fit(X, X[:k].copy())
    bad = int((labels != z[name + "_labels"]).sum())
    print(f"\n{name} (tensor-core path): {bad} of {labels.size} labels differ from sklearn")
    assert bad == 0
    assert np.abs(centres - z[name + "_centers"]).max() < 1e-4
    assert abs(inertia - float(z[name + "_inertia"])) <= 1e-5 * float(z[name + "_inertia"])
    assert n_iter == int(z[name + "_niter"])


def test_lloyd_tensor_core_path_on_off(monkeypatch):
    """the same fit through the stepwise float32 kernels (OFC_KMEANS_TC=0) and through the tensor-core path"""
    from opticalflowclustering_b200 import kmeans
    g = torch.Generator().manual_seed(5)
    cen = torch.rand((40, 96), generator=g) * 6
    X = (cen[torch.randint(0, 40, (30000,), generator=g)] + torch.randn((30000, 96), generator=g)).float().cuda()
    init = X[:40].double()
    monkeypatch.setenv("OFC_KMEANS_TC", "0")
    l0, c0, i0, n0 = kmeans.lloyd(X, init)
    monkeypatch.setenv("OFC_KMEANS_TC", "1")
    l1, c1, i1, n1 = kmeans.lloyd(X, init)
    l2, c2, i2, n2 = kmeans.lloyd(X, init)
    assert torch.equal(l1, l2) and torch.equal(c1, c2) and float(i1) == float(i2)          # deterministic
    assert int(n0) == int(n1)
    assert (l0 == l1).float().mean().item() > 0.9995
    assert (c0 - c1).abs().max().item() < 1e-4
    assert abs(float(i0) - float(i1)) <= 1e-6 * float(i0)


def test_lloyd_dense_full_size_properties():
    """BASELINE configs[4] scale (1 M x 64 float32 rows, k = 64) through the tensor-core path -- too large for the
    oracle, so size-independent properties: (a) two runs are bit-identical, (b) the labels are the arg-min of an
    independent float64 torch restatement on the final centres up to float32 near-ties, (c) run to a strict stop
    (tol = 0) the centres are the member means, (d) inertia equals the direct sum."""
    from opticalflowclustering_b200 import kmeans
    g = torch.Generator(device="cuda").manual_seed(4)
    N, D, K = 1_000_000, 64, 64
    cen = torch.rand((K, D), device="cuda", generator=g) * 8
    X = (cen[torch.randint(0, K, (N,), device="cuda", generator=g)] + torch.randn((N, D), device="cuda", generator=g)).float()
    init = X[:K].double()
    l1, c1, i1, n1 = kmeans.lloyd(X, init, tol=0.0)
    l2, c2, i2, n2 = kmeans.lloyd(X, init, tol=0.0)
    assert int(n1) < 300
    assert torch.equal(l1, l2) and torch.equal(c1, c2) and float(i1) == float(i2) and int(n1) == int(n2)
    Xd = X.double()
    d2 = (c1 ** 2).sum(1)[None, :] - 2.0 * Xd @ c1.T
    ref = d2.argmin(1).to(torch.int32)
    assert (ref != l1).sum().item() <= 20                                    # float32 near-ties only (of 1 M rows)
    sums = torch.zeros((K, D), dtype=torch.float64, device="cuda").index_add_(0, l1.long(), Xd)
    cnt = torch.bincount(l1.long(), minlength=K).double()
    assert ((sums / cnt[:, None]) - c1).abs().max().item() < 1e-5           # float32 centring / rounding of the centres
    direct = ((Xd - c1[l1.long()]) ** 2).sum().item()
    assert abs(direct - float(i1)) <= 1e-5 * direct
