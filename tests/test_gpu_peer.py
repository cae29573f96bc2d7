"""GPU test of the NVLink peer-memory exchange kernel (csrc/peer_exchange.cu) on ONE device: the `world` ranks are
`world` buffers of this process and `world` streams, so the flag protocol (publish / wait / pull in rank order / two
alternating slots / gated gather) runs exactly as between processes, minus the IPC mapping.  The multi-process run over
real NVLink is tools/dist_kmeans_check.py (profiles/r02k_dist_kmeans_*gpu.log)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

_vp = C.c_void_p


@pytest.fixture(scope="module")
def lib():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from opticalflowclustering_b200 import _lib
    return _lib.lib()


class _Ranks:
    def __init__(self, lib, world, region_bytes):
        self.lib, self.world = lib, world
        self.header = int(lib.ofc_peer_header_bytes())
        self.total = self.header + 3 * region_bytes
        self.ptrs = []
        for _ in range(world):
            p = _vp()
            assert lib.ofc_peer_alloc(C.c_size_t(self.total), C.byref(p), None) == 0
            self.ptrs.append(p)
        self.bufs = torch.tensor([p.value for p in self.ptrs], dtype=torch.int64, device="cuda")
        from opticalflowclustering_b200.peer import _RawDeviceBytes
        self.views = [torch.as_tensor(_RawDeviceBytes(p.value, self.total), device="cuda") for p in self.ptrs]
        self.streams = [torch.cuda.Stream() for _ in range(world)]
        self.region_bytes = region_bytes

    def offset(self, slot):
        return self.header + slot * self.region_bytes

    def write(self, rank, slot, f64, i64=None):
        """fill rank's region (on rank's stream, like the E/M kernels would)"""
        raw = np.concatenate([np.asarray(f64, np.float64).reshape(-1).view(np.uint8),
                              np.asarray(i64 if i64 is not None else [], np.int64).reshape(-1).view(np.uint8)])
        t = torch.from_numpy(raw).cuda()
        torch.cuda.current_stream().synchronize()
        with torch.cuda.stream(self.streams[rank]):
            o = self.offset(slot)
            self.views[rank][o:o + raw.nbytes].copy_(t)
        t.record_stream(self.streams[rank])

    def exchange(self, rank, slot, mode, n_f64, n_i64, out_f64, out_i64, gate=None):
        rc = self.lib.ofc_peer_exchange(_vp(self.bufs.data_ptr()), self.world, rank, C.c_size_t(self.offset(slot)), mode,
                                        C.c_int64(n_f64), C.c_int64(n_i64), _vp(out_f64.data_ptr() if out_f64 is not None else 0),
                                        _vp(out_i64.data_ptr() if out_i64 is not None else 0),
                                        _vp(gate.data_ptr() if gate is not None else 0), gate.numel() if gate is not None else 0,
                                        C.c_double(5.0), _vp(self.streams[rank].cuda_stream))
        assert rc == 0, self.lib.ofc_last_error().decode()

    def errors(self):
        torch.cuda.synchronize()
        out = []
        for p in self.ptrs:
            e = C.c_int(0)
            assert self.lib.ofc_peer_error(p, C.byref(e)) == 0
            out.append(e.value)
        return out

    def free(self):
        torch.cuda.synchronize()
        for p in self.ptrs:
            self.lib.ofc_peer_free(p)


@pytest.mark.parametrize("world,n_f64,n_i64", [(2, 32, 9), (4, 32, 9), (3, 1024 * 128, 1025)])
def test_peer_all_reduce_rank_order(lib, world, n_f64, n_i64):
    rng = np.random.default_rng(world * 7 + n_f64)
    R = _Ranks(lib, world, (n_f64 + n_i64) * 8 + 256 - ((n_f64 + n_i64) * 8) % 256)
    outs = [(torch.zeros(n_f64, dtype=torch.float64, device="cuda"), torch.zeros(n_i64, dtype=torch.int64, device="cuda"))
            for _ in range(world)]
    torch.cuda.synchronize()
    for it in range(6):                                    # six exchanges over the two alternating slots
        slot = it & 1
        f = rng.normal(0, 1e3, (world, n_f64))
        i = rng.integers(-2 ** 40, 2 ** 40, (world, n_i64))
        for r in range(world):
            R.write(r, slot, f[r], i[r])
        for r in range(world):
            R.exchange(r, slot, 0, n_f64, n_i64, outs[r][0], outs[r][1])
        assert R.errors() == [0] * world
        want = np.zeros(n_f64)
        for r in range(world):                             # rank order, like the kernel: bit-identical on every rank
            want = want + f[r]
        for r in range(world):
            assert (outs[r][0].cpu().numpy() == want).all()
            assert (outs[r][1].cpu().numpy() == i.sum(0)).all()
    R.free()


def test_peer_gather_gated(lib):
    world, n = 3, 16 * 7
    R = _Ranks(lib, world, 4096)
    rng = np.random.default_rng(3)
    outs = [torch.full((world, n), -1.0, dtype=torch.float64, device="cuda") for _ in range(world)]
    red = [(torch.zeros(4, dtype=torch.float64, device="cuda"), torch.zeros(2, dtype=torch.int64, device="cuda")) for _ in range(world)]
    full = torch.tensor([3, 1, 9], dtype=torch.int64, device="cuda")          # no empty cluster: the gather is skipped
    some_empty = torch.tensor([3, 0, 9], dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    pay = rng.normal(size=(world, n))
    for r in range(world):
        R.write(r, 2, pay[r])
    for r in range(world):
        R.exchange(r, 2, 1, n, 0, outs[r], None, gate=full)
    assert R.errors() == [0] * world
    for r in range(world):
        assert (outs[r].cpu().numpy() == -1.0).all()
    # an all-reduce, then the gated gather with an empty cluster, then an all-reduce again (the order of a Lloyd iteration)
    for r in range(world):
        R.write(r, 0, np.full(4, r + 1.0), [r, 1])
    for r in range(world):
        R.exchange(r, 0, 0, 4, 2, red[r][0], red[r][1])
    for r in range(world):
        R.exchange(r, 2, 1, n, 0, outs[r], None, gate=some_empty)
    for r in range(world):
        R.write(r, 1, np.full(4, 10.0 * (r + 1)), [5, r])
    for r in range(world):
        R.exchange(r, 1, 0, 4, 2, red[r][0], red[r][1])
    assert R.errors() == [0] * world
    for r in range(world):
        assert (outs[r].cpu().numpy() == pay).all()
        assert (red[r][0].cpu().numpy() == 60.0).all() and red[r][1].cpu().tolist() == [15, 3]
    R.free()


def test_peer_timeout_sets_error_instead_of_hanging(lib):
    """a rank whose peer never shows up gives up after the time limit and reports it"""
    R = _Ranks(lib, 2, 4096)
    out = torch.zeros(4, dtype=torch.float64, device="cuda")
    R.write(0, 0, np.ones(4))
    rc = lib.ofc_peer_exchange(_vp(R.bufs.data_ptr()), 2, 0, C.c_size_t(R.offset(0)), 0, C.c_int64(4), C.c_int64(0), _vp(out.data_ptr()),
                               _vp(0), _vp(0), 0, C.c_double(0.2), _vp(R.streams[0].cuda_stream))
    assert rc == 0
    assert R.errors() == [1, 0]
    R.free()
