"""ctypes wrapper of the CPU oracle (oracle/h264_oracle.c).  TEST INFRASTRUCTURE ONLY.

May be imported from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, nowhere else.  Never used by the product path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libh264_oracle.so")


class Params(C.Structure):
    """Mirror of vcpenc_params (include/vcpenc.h)."""
    _fields_ = [(n, C.c_int32) for n in (
        "width", "height", "fps_num", "fps_den", "codec", "gop", "rc_mode", "qp_i", "qp_p",
        "bitrate", "maxrate", "bufsize", "slices", "deblock_idc", "entropy", "in_fmt",
        "in_width", "in_height", "faststart", "effort", "debug", "first_gop", "drop_audio", "transform8x8", "hevc_subpel", "hevc_sao", "hevc_intra_modes")] + [("reserved", C.c_int32 * 5)]


class FrameInfo(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("size", C.c_uint32), ("is_idr", C.c_uint8),
                ("qp", C.c_uint8), ("pad", C.c_uint8 * 2)]


class Dump(C.Structure):
    _fields_ = [("mv_prepass", C.c_void_p), ("mv_final", C.c_void_p),
                ("mb_type", C.c_void_p), ("cbp", C.c_void_p)]


def make_params(width, height, fps=30, gop=60, qp_i=24, qp_p=26, slices=1, deblock_idc=0, **kw):
    p = Params()
    p.width, p.height = width, height
    p.fps_num, p.fps_den = fps, 1
    p.gop = gop
    p.qp_i, p.qp_p = qp_i, qp_p
    p.slices = slices
    p.deblock_idc = deblock_idc
    p.effort = 1            # vcpenc_default_params: medium (0 = the fast tiers: no quarter-sample step)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < max(os.path.getmtime(os.path.join(HERE, f)) for f in
                                             ("h264_oracle.c", "hevc_oracle.inc.c", "../video_codec_pipeline_b200/csrc/vcp_algo.h")):
        subprocess.check_call(["make", "-C", HERE, "-s"] + (["-B"] if force else []))
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_encode.argtypes = [C.POINTER(Params), C.c_void_p, C.c_int, C.c_void_p, C.c_size_t,
                                    C.POINTER(C.c_size_t), C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_encode.restype = C.c_int
    return _lib


def frame_bytes(w, h):
    return w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2)


def encode_hevc(params: Params, frames: np.ndarray):
    """HEVC oracle (oracle/hevc_oracle.inc.c): yuv420p frames -> dict(stream, info, recon)."""
    L = lib()
    L.hevc_orc_encode.argtypes = [C.POINTER(Params), C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                  C.c_void_p, C.c_void_p]
    L.hevc_orc_encode.restype = C.c_int
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    fb = frame_bytes(params.width, params.height)
    n = frames.size // fb
    cap = n * fb + (1 << 16)
    out = np.empty(cap, np.uint8)
    out_len = C.c_size_t(0)
    info = (FrameInfo * n)()
    recon = np.empty((n, fb), np.uint8)
    rc = L.hevc_orc_encode(C.byref(params), frames.ctypes.data, n, out.ctypes.data, cap, C.byref(out_len),
                           C.cast(info, C.c_void_p), recon.ctypes.data)
    if rc != 0:
        raise RuntimeError("hevc_orc_encode failed rc=%d" % rc)
    return {"stream": out[: out_len.value].tobytes(), "info": [(i.offset, i.size, i.is_idr, i.qp) for i in info], "recon": recon}


def in_frame_bytes(p) -> int:
    w = p.in_width if p.in_width > 0 else p.width
    h = p.in_height if p.in_height > 0 else p.height
    cw, ch = (w + 1) // 2, (h + 1) // 2
    return {0: w * h + 2 * cw * ch, 1: w * h + 2 * cw * ch, 2: 3 * w * h, 3: 3 * w * h,
            4: w * h + 2 * cw * h, 5: 3 * w * h}[p.in_fmt]


def encode(params: Params, frames: np.ndarray, want_recon=True, want_dump=False):
    """frames: uint8 array [n, input frame bytes] (params.in_fmt at the input size; yuv420p at
    the display size by default).

    Returns dict(stream=bytes, info=[(offset,size,is_idr,qp)], recon=np.ndarray|None, dump=dict|None).
    """
    L = lib()
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    fb = frame_bytes(params.width, params.height)
    n = frames.size // in_frame_bytes(params)
    assert frames.size == n * in_frame_bytes(params)
    cap = n * fb + (1 << 16)
    out = np.empty(cap, dtype=np.uint8)
    out_len = C.c_size_t(0)
    info = (FrameInfo * n)()
    recon = np.empty((n, fb), dtype=np.uint8) if want_recon else None
    mbw, mbh = (params.width + 15) // 16, (params.height + 15) // 16
    dump = None
    d = None
    if want_dump:
        dump = {
            "mv_prepass": np.zeros((n, mbh, mbw, 2), np.int16),
            "mv_final": np.zeros((n, mbh, mbw, 2), np.int16),
            "mb_type": np.zeros((n, mbh, mbw), np.uint8),
            "cbp": np.zeros((n, mbh, mbw), np.uint8),
        }
        d = Dump(*(dump[k].ctypes.data for k in ("mv_prepass", "mv_final", "mb_type", "cbp")))
    rc = L.orc_encode(C.byref(params), frames.ctypes.data, n, out.ctypes.data, cap, C.byref(out_len),
                      C.cast(info, C.c_void_p), recon.ctypes.data if want_recon else None,
                      C.byref(d) if d is not None else None)
    if rc != 0:
        raise RuntimeError("orc_encode failed rc=%d" % rc)
    return {
        "stream": out[: out_len.value].tobytes(),
        "info": [(i.offset, i.size, i.is_idr, i.qp) for i in info],
        "recon": recon,
        "dump": dump,
    }
