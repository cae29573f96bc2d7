"""GPU check of the tensor-core k-means kernels (tc_kmeans.cu): labels must be bit-identical to the
float32 CUDA-core E-step, sums/counts must match the generic M-step; prints timings."""
import ctypes as C
import sys
import time

import numpy as np
import torch

from opticalflowclustering_b200 import _lib

L = _lib.lib()
vp = C.c_void_p


def P(t):
    return vp(t.data_ptr()) if t is not None else vp(0)


def run(n, d, k, seed=0, timing=False, blobs=True, ties=False):
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(seed)
    if blobs:
        cen = torch.rand((k, d), device=dev, generator=g) * 8
        X = cen[torch.randint(0, k, (n,), device=dev, generator=g)] + torch.randn((n, d), device=dev, generator=g)
    else:
        X = torch.rand((n, d), device=dev, generator=g) * 255
    X = X.float().contiguous()
    mean = X.double().mean(0).float().double().contiguous()
    Xh = torch.empty_like(X)
    Xl = torch.empty_like(X)
    xnorm = torch.empty(n, dtype=torch.float32, device=dev)
    st = vp(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.ofc_kmeans_tc_prepare(P(X), P(mean), n, d, P(Xh), P(Xl), P(xnorm), st))
    Xc = X - mean.float()
    assert torch.equal(Xh + Xl, Xc), "prepare: hi + lo is not the centred row"
    assert int((Xh.view(torch.int32) & 0x1fff).abs().max().item()) == 0, "prepare: hi part has low bits"
    centres = Xc[:k].double().contiguous()                       # init = first k rows (centred)
    if ties:                                                     # duplicated / nearly duplicated centres
        h = k // 2
        centres[h:2 * h] = centres[:h]
        centres[h:h + h // 2] += 1e-6 * torch.randn((h // 2, d), device=dev, generator=g, dtype=torch.float64)
        centres = centres.float().double().contiguous()
    wsb = int(L.ofc_kmeans_tc_workspace_bytes(n, d, k))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    labels = torch.full((n,), -1, dtype=torch.int32, device=dev)
    prev = torch.full((n,), -1, dtype=torch.int32, device=dev)
    n_changed = torch.zeros(1, dtype=torch.int64, device=dev)
    inertia = torch.zeros(1, dtype=torch.float64, device=dev)
    nre = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(L.ofc_kmeans_tc_assign(P(Xh), P(Xl), P(xnorm), n, d, k, P(centres), P(labels), P(prev), P(n_changed), P(inertia),
                                      P(nre), P(ws), wsb, st))
    torch.cuda.synchronize()
    # reference: generic float32 kernel on the centred rows
    wsb2 = int(L.ofc_kmeans_workspace_bytes(1, n, d, k))
    ws2 = torch.empty(max(wsb2, 256), dtype=torch.uint8, device=dev)
    lab2 = torch.empty(n, dtype=torch.int32, device=dev)
    in2 = torch.zeros(1, dtype=torch.float64, device=dev)
    t0 = time.time()
    _lib.check(L.ofc_kmeans_assign(P(Xc), 1, 1, n, d, k, vp(0), P(centres), P(lab2), vp(0), vp(0), P(in2), vp(0), vp(0),
                                   P(ws2), max(wsb2, 256), st))
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    mism = int((labels != lab2).sum().item())
    # float64 ground truth for information
    print(f"n={n} d={d} k={k}: label mismatches vs float32 CUDA-core path {mism}, re-evaluated rows {int(nre.item())} "
          f"({100.0 * int(nre.item()) / n:.2f} %), n_changed {int(n_changed.item())}, inertia rel diff "
          f"{abs(inertia.item() - in2.item()) / max(in2.item(), 1e-30):.2e}, generic {t_gen * 1e3:.1f} ms", flush=True)
    ok = mism == 0 and int(n_changed.item()) == n
    # M-step
    sums = torch.empty((k, d), dtype=torch.float64, device=dev)
    counts = torch.empty(k, dtype=torch.int64, device=dev)
    _lib.check(L.ofc_kmeans_tc_sums(P(Xh), P(Xl), n, d, k, P(labels), P(sums), P(counts), P(ws), wsb, st))
    torch.cuda.synchronize()
    cnt_ref = torch.bincount(labels.long(), minlength=k)
    sums_ref = torch.zeros((k, d), dtype=torch.float64, device=dev).index_add_(0, labels.long(), Xc.double())
    cnt_bad = int((counts != cnt_ref).sum().item())
    err = float((sums - sums_ref).abs().max().item())
    scale = float(sums_ref.abs().max().item())
    print(f"   M-step: count mismatches {cnt_bad}, max |sum diff| {err:.3e} (scale {scale:.3e})", flush=True)
    ok = ok and cnt_bad == 0 and err <= 1e-9 * max(scale, 1.0)
    if timing:
        for name, fn in (("tc_assign", lambda: L.ofc_kmeans_tc_assign(P(Xh), P(Xl), P(xnorm), n, d, k, P(centres), P(labels), vp(0), vp(0),
                                                                     vp(0), vp(0), P(ws), wsb, st)),
                         ("tc_sums", lambda: L.ofc_kmeans_tc_sums(P(Xh), P(Xl), n, d, k, P(labels), P(sums), P(counts), P(ws), wsb, st))):
            for _ in range(2):
                fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                fn()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            extra = f", {2.0 * n * k * d / ms / 1e9:.1f} TFLOP/s" if name == "tc_assign" else f", {n * d * 4 / ms / 1e6:.0f} GB/s"
            print(f"   {name}: {ms:.3f} ms{extra}", flush=True)
    return ok


def run_u8(n, d, k, seed=0, timing=False, ties=False):
    """uint8 rows: tensor-core filter + float64 re-evaluation must give the labels of the float64 CUDA-core E-step"""
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(seed)
    cen = torch.rand((k, d), device=dev, generator=g) * 200 + 20
    X = (cen[torch.randint(0, k, (n,), device=dev, generator=g)] + 12 * torch.randn((n, d), device=dev, generator=g))
    X = X.round().clamp(0, 255).to(torch.uint8).contiguous()
    mean = X.double().mean(0).contiguous()
    Xh = torch.empty((n, d), dtype=torch.float32, device=dev)
    Xl = torch.empty_like(Xh)
    xnorm = torch.empty(n, dtype=torch.float32, device=dev)
    st = vp(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.ofc_kmeans_tc_prepare_u8(P(X), P(mean), n, d, P(Xh), P(Xl), P(xnorm), st))
    assert torch.equal(Xh + Xl, (X.double() - mean).float()), "prepare_u8: hi + lo is not float32(x - mean)"
    centres = (X[:k].double() - mean).contiguous()
    if ties:
        h = k // 2
        centres[h:2 * h] = centres[:h]
        centres[h:h + h // 2] += 1e-9 * torch.randn((h // 2, d), device=dev, generator=g, dtype=torch.float64)
    wsb = int(L.ofc_kmeans_tc_workspace_bytes(n, d, k))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    labels = torch.full((n,), -1, dtype=torch.int32, device=dev)
    prev = torch.full((n,), -1, dtype=torch.int32, device=dev)
    n_changed = torch.zeros(1, dtype=torch.int64, device=dev)
    nre = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(L.ofc_kmeans_tc_assign_u8(P(X), P(mean), P(Xh), P(Xl), P(xnorm), n, d, k, P(centres), P(labels), P(prev), P(n_changed),
                                         P(nre), P(ws), wsb, st))
    torch.cuda.synchronize()
    wsb2 = int(L.ofc_kmeans_workspace_bytes(1, n, d, k))
    ws2 = torch.empty(max(wsb2, 256), dtype=torch.uint8, device=dev)
    lab2 = torch.empty(n, dtype=torch.int32, device=dev)
    t0 = time.time()
    _lib.check(L.ofc_kmeans_assign(P(X), 0, 1, n, d, k, P(mean), P(centres), P(lab2), vp(0), vp(0), vp(0), vp(0), vp(0),
                                   P(ws2), max(wsb2, 256), st))
    torch.cuda.synchronize()
    t_ref = time.time() - t0
    mism = int((labels != lab2).sum().item())
    print(f"uint8 n={n} d={d} k={k}: label mismatches vs float64 CUDA-core path {mism}, re-evaluated rows {int(nre.item())} "
          f"({100.0 * int(nre.item()) / n:.2f} %), float64 kernel {t_ref * 1e3:.1f} ms", flush=True)
    ok = mism == 0 and int(n_changed.item()) == n
    sums = torch.empty((k, d), dtype=torch.float64, device=dev)
    counts = torch.empty(k, dtype=torch.int64, device=dev)
    _lib.check(L.ofc_kmeans_tc_sums_u8(P(X), n, d, k, P(labels), P(sums), P(counts), P(ws), wsb, st))
    torch.cuda.synchronize()
    cnt_ref = torch.bincount(labels.long(), minlength=k)
    sums_ref = torch.zeros((k, d), dtype=torch.float64, device=dev).index_add_(0, labels.long(), X.double())
    ok = ok and torch.equal(counts, cnt_ref) and torch.equal(sums, sums_ref)
    print(f"   M-step: counts exact {torch.equal(counts, cnt_ref)}, integer sums exact {torch.equal(sums, sums_ref)}", flush=True)
    if timing:
        fn = lambda: L.ofc_kmeans_tc_assign_u8(P(X), P(mean), P(Xh), P(Xl), P(xnorm), n, d, k, P(centres), P(labels), vp(0), vp(0), vp(0),
                                               P(ws), wsb, st)
        for _ in range(2):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"   tc_assign_u8: {ms:.3f} ms, {2.0 * n * k * d / ms / 1e9:.1f} TFLOP/s", flush=True)
    return ok


if __name__ == "__main__":
    if len(sys.argv) > 4 and sys.argv[1] == "one":
        sys.exit(0 if run(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), timing=True) else 1)
    shapes = [(1000, 32, 8), (4096, 64, 64), (5000, 100, 40), (20000, 128, 300), (100000, 64, 256), (50000, 512, 1024),
              (3001, 36, 2)]
    ok = True
    for (n, d, k) in shapes:
        ok = run(n, d, k) and ok
    ok = run(20000, 64, 64, blobs=False) and ok
    ok = run(30000, 96, 128, ties=True) and ok
    ok = run(30000, 256, 512, ties=True, blobs=False) and ok
    for (n, d, k) in [(3000, 36, 8), (20000, 352, 8), (50000, 128, 256), (30000, 352, 1024)]:
        ok = run_u8(n, d, k) and ok
    ok = run_u8(20000, 64, 128, ties=True) and ok
    if len(sys.argv) > 1 and sys.argv[1] == "time":
        for (n, d, k) in [(1000000, 64, 256), (1000000, 128, 1024), (1000000, 512, 256), (200000, 2048, 1024)]:
            ok = run(n, d, k, timing=True) and ok
        for (n, d, k) in [(1000000, 352, 256), (1000000, 128, 1024)]:
            ok = run_u8(n, d, k, timing=True) and ok
    print("TC_CHECK", "OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)
