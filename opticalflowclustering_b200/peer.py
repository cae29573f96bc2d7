"""NVLink peer-memory exchange for the row-sharded k-means (``kmeans.lloyd(group=...)`` on the GPUs of ONE node).

Per Lloyd iteration the ranks exchange the per-cluster sums / counts / label changes (264 bytes for k = 8 x 4 features,
1 MB for k = 1024 x 128) and, when a cluster is empty, their far-point lists.  Through ``torch.distributed`` that is three
library collectives issued from Python per iteration -- more host time than the kernels of the iteration take.  Here every
rank owns one device buffer (``ofc_peer_alloc``), maps the others' through CUDA IPC, and the exchange is ONE kernel
(``ofc_peer_exchange``, csrc/peer_exchange.cu): flags + pull over NVLink, summed in rank order, so every rank holds
bit-identical values and the whole iteration is plain kernel launches (capturable in a CUDA graph).

torch.distributed is used once per (group, size): to pass the 64-byte IPC handles around and to agree that every rank
managed to map every peer.  If any rank cannot (different hosts, no peer access, IPC forbidden), all ranks fall back to
the NCCL collectives -- ``PeerExchange.create`` returns None everywhere.
"""
from __future__ import annotations

import ctypes as C
import os
import socket

import torch

from . import _lib

_vp = C.c_void_p


class _RawDeviceBytes:
    """``__cuda_array_interface__`` view of a raw device allocation (so the regions can be handed to the kernels and to
    torch ops as ordinary tensors)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerExchange:
    """One symmetric buffer per rank: header | slot 0 | slot 1 | gather region."""

    _cache: dict = {}

    def __init__(self, lib, group, device, slot_bytes: int, gather_bytes: int):
        self.lib, self.group, self.device = lib, group, device
        self.header = int(lib.ofc_peer_header_bytes())
        self.slot_bytes = (int(slot_bytes) + 255) // 256 * 256
        self.gather_bytes = (int(gather_bytes) + 255) // 256 * 256
        self.total = self.header + 2 * self.slot_bytes + self.gather_bytes
        self.own = _vp()
        self.peers: list = []
        self.opened: list = []
        self.ok = False

    # region offsets (bytes from the start of the buffer)
    def slot_offset(self, s: int) -> int:
        return self.header + s * self.slot_bytes

    def gather_offset(self) -> int:
        return self.header + 2 * self.slot_bytes

    @classmethod
    def create(cls, group, device, slot_bytes: int, gather_bytes: int):
        """The exchange object of this process group (cached; grown when a larger fit comes along), or None when the
        ranks cannot map each other's memory.  Collective: every rank of the group must call it with the same sizes."""
        import torch.distributed as dist
        if device.type != "cuda" or os.environ.get("OFC_KMEANS_PEER", "1") == "0" or dist.get_backend(group) != "nccl":
            return None
        world = dist.get_world_size(group)
        if world < 2 or world > 16:
            return None
        key = (id(group) if group is not None else 0, device.index)
        have = cls._cache.get(key)
        if have is not None:
            if have is False:
                return None
            if have.slot_bytes >= slot_bytes and have.gather_bytes >= gather_bytes:
                return have
            # a larger fit: every rank must be done with the old buffers before anybody unmaps or frees them
            torch.cuda.synchronize(device)
            dist.barrier(group=group)
            have.release()
        lib = _lib.lib()
        px = cls(lib, group, device, max(slot_bytes, 1 << 16), max(gather_bytes, 1 << 16))
        px._setup(dist, world)
        cls._cache[key] = px if px.ok else False
        return px if px.ok else None

    def _setup(self, dist, world):
        lib = self.lib
        rank = dist.get_rank(self.group)
        self.world, self.rank = world, rank
        handle = (C.c_ubyte * 64)()
        good = 1
        with torch.cuda.device(self.device):
            rc = lib.ofc_peer_alloc(C.c_size_t(self.total), C.byref(self.own), handle)
        if rc != 0:
            good = 0
        # handles + host names travel as one byte tensor per rank (NCCL all-gather: no pickling on the way)
        host = socket.gethostname().encode()[:63]
        msg = torch.zeros(130, dtype=torch.uint8)
        msg[:64] = torch.tensor(list(bytes(handle)), dtype=torch.uint8)
        msg[64] = good
        msg[65] = len(host)
        msg[66:66 + len(host)] = torch.tensor(list(host), dtype=torch.uint8)
        mine = msg.to(self.device)
        everyone = torch.empty((world, 130), dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(everyone, mine, group=self.group)
        everyone = everyone.cpu()
        same_host = all(bytes(everyone[r, 66:66 + int(everyone[r, 65])].tolist()) == host for r in range(world))
        good = good and same_host and all(int(everyone[r, 64]) for r in range(world))
        self.peers = [None] * world
        if good:
            with torch.cuda.device(self.device):
                for r in range(world):
                    if r == rank:
                        self.peers[r] = self.own.value
                        continue
                    p = _vp()
                    h = (C.c_ubyte * 64)(*everyone[r, :64].tolist())
                    if lib.ofc_peer_open(h, C.byref(p)) != 0:
                        good = 0
                        break
                    self.opened.append(p)
                    self.peers[r] = p.value
        agree = torch.tensor([1 if good else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=self.group)
        if int(agree.item()) == 0:
            self.release()
            return
        self.bufs = torch.tensor(self.peers, dtype=torch.int64, device=self.device)      # device array of the mapped pointers
        self.raw = torch.as_tensor(_RawDeviceBytes(self.own.value, self.total), device=self.device)
        # nobody may publish into a buffer before every rank has mapped it and zeroed its flags (ofc_peer_alloc)
        dist.barrier(group=self.group)
        self.ok = True

    def region(self, offset: int, nbytes: int) -> torch.Tensor:
        """uint8 view of this rank's own buffer"""
        return self.raw[offset:offset + nbytes]

    def exchange(self, offset: int, mode: int, n_f64: int, n_i64: int, out_f64, out_i64, gate, stream, timeout_s: float = 10.0):
        rc = self.lib.ofc_peer_exchange(_vp(self.bufs.data_ptr()), self.world, self.rank, C.c_size_t(offset), int(mode),
                                        C.c_int64(n_f64), C.c_int64(n_i64), _vp(out_f64.data_ptr() if out_f64 is not None else 0),
                                        _vp(out_i64.data_ptr() if out_i64 is not None else 0),
                                        _vp(gate.data_ptr() if gate is not None else 0), int(gate.numel()) if gate is not None else 0,
                                        C.c_double(timeout_s), stream)
        if rc != 0:
            raise _lib.OfcError(self.lib.ofc_last_error().decode())

    def check(self):
        """after a stream synchronise: did a wait for a peer time out?"""
        err = C.c_int(0)
        self.lib.ofc_peer_error(self.own, C.byref(err))
        if err.value:
            raise _lib.OfcError("peer exchange: a rank did not publish its data within the time limit")

    def release(self):
        self.ok = False
        self.raw = self.bufs = None
        for p in self.opened:
            self.lib.ofc_peer_close(p)
        self.opened = []
        if self.own.value:
            self.lib.ofc_peer_free(self.own)
            self.own = _vp()
