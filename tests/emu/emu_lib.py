"""ctypes access to the host-side debug emulation of libofc (tests only).

Builds tests/emu/libofc_emu.so from the very same .cu sources with g++
(-DOFC_EMULATE, tests/emu/cuda_emu.h) and exposes numpy-in / numpy-out helpers.
The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "opticalflowclustering_b200", "csrc")
SO = os.path.join(HERE, "libofc_emu.so")


def build(force: bool = False) -> str:
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    srcs = [s for s in srcs if not os.path.basename(s).startswith("tc_")]     # tcgen05 kernels: GPU only
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "cuda_emu.h"),
                                                            os.path.join(ROOT, "include", "ofc.h")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return SO
    cmd = ["g++", "-O2", "-std=c++17", "-DOFC_EMULATE", "-include", os.path.join(HERE, "cuda_emu.h"),
           "-x", "c++", "-fPIC", "-shared", "-o", SO] + srcs
    subprocess.run(cmd, check=True)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.ofc_last_error.restype = C.c_char_p
        _lib.ofc_flow_plan_workspace_bytes.restype = C.c_size_t
        _lib.ofc_flow_plan_workspace_bytes.argtypes = [C.c_void_p]
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


def _check(rc):
    if rc != 0:
        raise RuntimeError(f"libofc_emu rc={rc}: {lib().ofc_last_error().decode()}")


class Plan:
    def __init__(self, W, H, max_frames=2, pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5,
                 poly_sigma=1.2, flags=0):
        L = lib()
        self.ptr = C.c_void_p()
        _check(L.ofc_flow_plan_create(C.byref(self.ptr), W, H, max_frames, C.c_double(pyr_scale), levels, winsize,
                                      iterations, poly_n, C.c_double(poly_sigma), flags))
        self.W, self.H, self.max_frames = W, H, max_frames
        self.ws_bytes = L.ofc_flow_plan_workspace_bytes(self.ptr)
        raw = np.zeros(self.ws_bytes + 256, np.uint8)
        off = (-raw.ctypes.data) % 256
        self.ws = raw[off:off + self.ws_bytes]
        self.n_levels = L.ofc_flow_plan_num_levels(self.ptr)
        _check(L.ofc_flow_plan_keep_intermediates(self.ptr, 1))

    def level_size(self, l):
        w, h = C.c_int(), C.c_int()
        _check(lib().ofc_flow_plan_level_size(self.ptr, l, C.byref(w), C.byref(h)))
        return w.value, h.value

    def buffer(self, level, kind, frame=0):
        off, st = C.c_size_t(), C.c_size_t()
        _check(lib().ofc_flow_plan_buffer(self.ptr, level, kind, C.byref(off), C.byref(st)))
        w, h = self.level_size(level)
        ch = {0: 1, 1: 4, 2: 1, 3: 2, 4: 2}[kind]
        b = self.ws[off.value + frame * st.value: off.value + (frame + 1) * st.value]
        return b.view(np.float32).reshape(h, w, ch)

    def sequence(self, gray, want_minmax=False):
        gray = np.ascontiguousarray(gray, np.uint8)
        n = gray.shape[0]
        flow = np.zeros((n - 1, self.H, self.W, 2), np.float32)
        mm = np.zeros((n - 1, 2), np.uint32) if want_minmax else None
        _check(lib().ofc_farneback_sequence(self.ptr, _p(gray), n, _p(flow), _p(mm), _p(self.ws),
                                            C.c_size_t(self.ws_bytes), C.c_void_p(0)))
        return (flow, mm) if want_minmax else flow

    def pair_init(self, prev, nxt, init_flow):
        """cv2 flag OPTFLOW_USE_INITIAL_FLOW: init_flow f32 [H,W,2]"""
        g = np.ascontiguousarray(np.stack([prev, nxt]), np.uint8)
        init = np.ascontiguousarray(init_flow, np.float32)
        flow = np.zeros((self.H, self.W, 2), np.float32)
        _check(lib().ofc_farneback_pair_init(self.ptr, _p(g[0]), _p(g[1]), _p(init), _p(flow), C.c_void_p(0), _p(self.ws),
                                             C.c_size_t(self.ws_bytes), C.c_void_p(0)))
        return flow

    def stream_begin(self, gray0):
        g = np.ascontiguousarray(gray0, np.uint8)
        _check(lib().ofc_farneback_stream_begin(self.ptr, _p(g), _p(self.ws), C.c_size_t(self.ws_bytes), C.c_void_p(0)))

    def stream_next(self, gray):
        g = np.ascontiguousarray(gray, np.uint8)
        flow = np.zeros((self.H, self.W, 2), np.float32)
        _check(lib().ofc_farneback_stream_next(self.ptr, _p(g), _p(flow), C.c_void_p(0), _p(self.ws), C.c_size_t(self.ws_bytes),
                                               C.c_void_p(0)))
        return flow

    def __del__(self):
        try:
            lib().ofc_flow_plan_destroy(self.ptr)
        except Exception:
            pass


def bgr2gray(bgr):
    bgr = np.ascontiguousarray(bgr, np.uint8)
    gray = np.zeros(bgr.shape[:-1], np.uint8)
    _check(lib().ofc_bgr2gray(_p(bgr), _p(gray), C.c_int64(gray.size), C.c_void_p(0)))
    return gray


def flow_to_bgr(flow, minmax=None):
    flow = np.ascontiguousarray(flow, np.float32)
    n = flow.shape[0]
    npx = flow.shape[1] * flow.shape[2]
    if minmax is None:
        minmax = np.zeros((n, 2), np.uint32)
        _check(lib().ofc_flow_minmax(_p(flow), n, C.c_int64(npx), _p(minmax), C.c_void_p(0)))
    bgr = np.zeros(flow.shape[:3] + (3,), np.uint8)
    mag = np.zeros(n, np.float64)
    _check(lib().ofc_flow_to_bgr(_p(flow), n, flow.shape[1], flow.shape[2], _p(minmax), _p(bgr), _p(mag), C.c_void_p(0)))
    return bgr, mag, minmax


def grid_cells(bgr, rows, cols, draw_lines=1, threshold=30):
    bgr = np.ascontiguousarray(bgr, np.uint8)
    n, H, W, _ = bgr.shape
    cells = rows * cols
    out = dict(avg_bgr=np.zeros((n, cells, 3), np.uint8), avg_hue=np.zeros((n, cells), np.uint8),
               km_centre=np.zeros((n, cells, 4), np.uint8), km_hue=np.zeros((n, cells), np.uint8),
               km_sums=np.zeros((n, cells, 4), np.uint32))
    _check(lib().ofc_grid_cells(_p(bgr), n, H, W, rows, cols, draw_lines, threshold, _p(out["avg_bgr"]),
                                _p(out["avg_hue"]), _p(out["km_centre"]), _p(out["km_hue"]), _p(out["km_sums"]),
                                C.c_void_p(0)))
    return out


def draw_grid(bgr, rows, cols):
    bgr = np.ascontiguousarray(bgr, np.uint8).copy()
    n, H, W, _ = bgr.shape
    _check(lib().ofc_draw_grid(_p(bgr), n, H, W, rows, cols, C.c_void_p(0)))
    return bgr


def flow_to_bgr_grid(flow, rows, cols, minmax=None, draw_lines=1, threshold=30):
    """the fused visualisation + grid kernel; returns (bgr, mag_sum, grid outputs)"""
    flow = np.ascontiguousarray(flow, np.float32)
    n, H, W = flow.shape[:3]
    if minmax is None:
        minmax = np.zeros((n, 2), np.uint32)
        _check(lib().ofc_flow_minmax(_p(flow), n, C.c_int64(H * W), _p(minmax), C.c_void_p(0)))
    bgr = np.zeros((n, H, W, 3), np.uint8)
    mag = np.zeros(n, np.float64)
    cells = rows * cols
    out = dict(avg_bgr=np.zeros((n, cells, 3), np.uint8), avg_hue=np.zeros((n, cells), np.uint8),
               km_centre=np.zeros((n, cells, 4), np.uint8), km_hue=np.zeros((n, cells), np.uint8))
    _check(lib().ofc_flow_to_bgr_grid(_p(flow), n, H, W, _p(minmax), _p(bgr), _p(mag), rows, cols, draw_lines, threshold,
                                      _p(out["avg_bgr"]), _p(out["avg_hue"]), _p(out["km_centre"]), _p(out["km_hue"]),
                                      C.c_void_p(0)))
    return bgr, mag, out


def flow_to_hsv(flow, minmax=None):
    """BGR visualisation + the reference's `mask` (H, 255, V)"""
    flow = np.ascontiguousarray(flow, np.float32)
    n, H, W = flow.shape[:3]
    if minmax is None:
        minmax = np.zeros((n, 2), np.uint32)
        _check(lib().ofc_flow_minmax(_p(flow), n, C.c_int64(H * W), _p(minmax), C.c_void_p(0)))
    bgr = np.zeros((n, H, W, 3), np.uint8)
    hsv = np.zeros((n, H, W, 3), np.uint8)
    _check(lib().ofc_flow_to_hsv(_p(flow), n, H, W, _p(minmax), _p(bgr), _p(hsv), C.c_void_p(0)))
    return bgr, hsv
