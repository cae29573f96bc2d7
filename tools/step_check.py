"""Short driver for profiling the stepwise k-means kernels: fused uint8 step (64 M x 4, k = 8) and the medium
E-step (uint8 200 k x 350, k = 8)."""
import torch

from opticalflowclustering_b200 import kmeans as km

dev = torch.device("cuda")
for (N, D, K) in [(64_000_000, 4, 8), (200_000, 350, 8)]:
    X = torch.randint(0, 180, (N, D), device=dev, dtype=torch.uint8)
    st = km.LloydState(km._Ctx(dev), X.unsqueeze(0).contiguous(), K)
    centres = X[:K].double().unsqueeze(0).contiguous()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fused = st.step_supported()
    for it in range(4):
        if it == 1:
            e0.record()
        if fused:
            st.step(None, centres, st.labels[it & 1], st.labels[(it & 1) ^ 1], st.n_changed, st.sums, st.counts)
        else:
            st.assign(None, centres, st.labels[0])
    e1.record()
    torch.cuda.synchronize()
    print(f"N={N} D={D} k={K} fused={fused}: {e0.elapsed_time(e1) / 3:.3f} ms per {'step' if fused else 'E-step'}")
