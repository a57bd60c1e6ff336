"""ctypes binding of libvcpenc.so (include/vcpenc.h) — the host-side mirror of the
reference's executor boundary.

The reference's boundary is two Go functions (`runFFmpegWithTimeout`,
/root/reference/cmd/consumer.go:370-394, and `verifyOutputFile`, :396-419).  Go is not in
this image, so the host mirror above the C-ABI is Python: `transcode()` / `verify()` keep the
reference's argument meaning and error behaviour; `encode_frames()` and `Session` expose the
in-memory core for tests and bench.py.  There is no CPU fallback: every encode entry point
raises if the CUDA library is missing or no device is visible.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

# more hardware queues than the default 8: the encoder overlaps ~45 streams (see encoder.cu); must be
# in the environment before the CUDA context exists, so also before torch touches the device
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libvcpenc.so")

K_NAMES = ["csc", "me_prepass", "me_refine", "p_recon", "i_recon", "mbinfo", "deblock", "pad",
           "cavlc_count", "cavlc_scan", "cavlc_write_pack", "rc", "hpel", "cabac_bins", "cabac_code"]

ERR_NAMES = {0: "OK", 1: "ARGS", 2: "IO", 3: "FORMAT", 4: "NODEVICE", 5: "CUDA", 6: "CANCELLED",
             7: "TIMEOUT", 8: "NOTENCODE", 9: "AUDIO", 10: "VERIFY", 11: "OVERFLOW", 12: "INTERNAL", 13: "UNSUPPORTED"}


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "width", "height", "fps_num", "fps_den", "codec", "gop", "rc_mode", "qp_i", "qp_p",
        "bitrate", "maxrate", "bufsize", "slices", "deblock_idc", "entropy", "in_fmt",
        "in_width", "in_height", "faststart", "effort", "debug", "first_gop", "drop_audio", "transform8x8", "hevc_subpel", "hevc_sao", "hevc_intra_modes", "audio_bitrate")] + [("reserved", C.c_int32 * 4)]


class FrameInfo(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("size", C.c_uint32), ("is_idr", C.c_uint8),
                ("qp", C.c_uint8), ("pad", C.c_uint8 * 2)]


class KernelStat(C.Structure):
    _fields_ = [("ms", C.c_double), ("launches", C.c_uint64)]


class VcpencError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("vcpenc error %d (%s): %s" % (code, ERR_NAMES.get(code, "?"), msg))
        self.code = code
        self.msg = msg


_lib = None


def lib():
    """Load libvcpenc.so; fails loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libvcpenc.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "or `make -C video_codec_pipeline_b200/csrc`")
        L = C.CDLL(LIB_PATH)
        vp, cp, sz = C.c_void_p, C.c_char_p, C.c_size_t
        L.vcpenc_version.restype = cp
        L.vcpenc_device_count.restype = C.c_int
        L.vcpenc_default_params.argtypes = [C.POINTER(Params)]
        L.vcpenc_encode_frames.argtypes = [C.POINTER(Params), C.c_int, vp, C.c_int, vp, sz, C.POINTER(sz), vp, vp,
                                           vp, cp, sz]
        L.vcpenc_session_create.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.POINTER(vp), cp, sz]
        L.vcpenc_session_upload.argtypes = [vp, vp, C.c_int, cp, sz]
        L.vcpenc_session_upload_async.argtypes = [vp, vp, C.c_int, cp, sz]
        L.vcpenc_session_upload_device.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_float), cp, sz]
        L.vcpenc_session_launch_count.argtypes = [vp]
        L.vcpenc_session_launch_count.restype = C.c_uint64
        L.vcpenc_session_set_first_gop.argtypes = [vp, C.c_int]
        L.vcpenc_host_alloc.argtypes = [sz]
        L.vcpenc_host_alloc.restype = vp
        L.vcpenc_host_free.argtypes = [vp]
        L.vcpenc_session_encode.argtypes = [vp, C.POINTER(C.c_float), cp, sz]
        L.vcpenc_session_download.argtypes = [vp, vp, sz, C.POINTER(sz), vp, vp, cp, sz]
        L.vcpenc_session_profile.argtypes = [vp, C.c_int]
        L.vcpenc_session_kernel_stats.argtypes = [vp, vp]
        L.vcpenc_session_debug_mbs.argtypes = [vp, vp, vp, vp, vp]
        L.vcpenc_session_destroy.argtypes = [vp]
        for name, args in (("vcpenc_parse_args", [C.c_int, C.POINTER(cp), C.POINTER(Params), cp, sz]),
                           ("vcpenc_transcode", [cp, cp, C.c_int, C.POINTER(cp), C.c_int, vp, cp, sz]),
                           ("vcpenc_verify", [cp, cp, sz]),
                           ("vcpenc_mux_mp4", [C.POINTER(Params), vp, sz, vp, C.c_int, cp, cp, sz])):
            if hasattr(L, name):
                getattr(L, name).argtypes = args
        _lib = L
    return _lib


def default_params(width, height, **kw) -> Params:
    p = Params()
    lib().vcpenc_default_params(C.byref(p))
    p.width, p.height = width, height
    for k, v in kw.items():
        if k == "fps":
            p.fps_num, p.fps_den = v, 1
        else:
            setattr(p, k, v)
    return p


def frame_bytes(w, h):
    return w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2)


def in_frame_bytes(p) -> int:
    """Bytes of one INPUT frame (p.in_fmt at p.in_width x p.in_height; vcp_in_frame_bytes)."""
    w = p.in_width if p.in_width > 0 else p.width
    h = p.in_height if p.in_height > 0 else p.height
    cw, ch = (w + 1) // 2, (h + 1) // 2
    return {0: w * h + 2 * cw * ch, 1: w * h + 2 * cw * ch, 2: 3 * w * h, 3: 3 * w * h,
            4: w * h + 2 * cw * h, 5: 3 * w * h}[p.in_fmt]


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a  # raw integer address (e.g. torch pinned tensor .data_ptr())


def device_count() -> int:
    return lib().vcpenc_device_count()


def version() -> str:
    return lib().vcpenc_version().decode()


def encode_frames(params: Params, frames, nframes=None, device=0, want_recon=False, out=None):
    """Host buffers in -> Annex-B out, through vcpenc_encode_frames (H2D + kernels + D2H)."""
    L = lib()
    fb = frame_bytes(params.width, params.height)
    if isinstance(frames, np.ndarray):
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        nframes = frames.size // in_frame_bytes(params)
    cap = nframes * fb + (1 << 20)
    if out is None:
        out = np.empty(cap, np.uint8)
    out_len = C.c_size_t(0)
    info = (FrameInfo * nframes)()
    recon = np.empty((nframes, fb), np.uint8) if want_recon else None
    err = C.create_string_buffer(512)
    rc = L.vcpenc_encode_frames(C.byref(params), device, _ptr(frames), nframes, _ptr(out), out.size,
                                C.byref(out_len), C.cast(info, C.c_void_p), _ptr(recon), None, err, 512)
    if rc:
        raise VcpencError(rc, err.value.decode(errors="replace"))
    return {"stream": out[: out_len.value], "info": [(i.offset, i.size, i.is_idr, i.qp) for i in info],
            "recon": recon}


class Session:
    """Frames stay resident in HBM; encode() can be timed on the device."""

    def __init__(self, params: Params, max_frames: int, device: int = 0):
        self.L = lib()
        self.params = params
        self.max_frames = max_frames
        self.h = C.c_void_p(None)
        self.err = C.create_string_buffer(512)
        rc = self.L.vcpenc_session_create(C.byref(params), device, max_frames, C.byref(self.h), self.err, 512)
        if rc:
            raise VcpencError(rc, self.err.value.decode(errors="replace"))
        self.nframes = 0
        self.fb = frame_bytes(params.width, params.height)
        self.mbw, self.mbh = (params.width + 15) // 16, (params.height + 15) // 16

    def _ck(self, rc):
        if rc:
            raise VcpencError(rc, self.err.value.decode(errors="replace"))

    def upload(self, frames, nframes=None, wait=True):
        """Host frames -> device planes.  wait=False: return once the copies are queued; the following encode()
        starts each GOP group when its frames have landed (frames must stay alive until encode() returns)."""
        if isinstance(frames, np.ndarray):
            frames = np.ascontiguousarray(frames, dtype=np.uint8)
            nframes = frames.size // in_frame_bytes(self.params)
        self._keep = frames
        fn = self.L.vcpenc_session_upload if wait else self.L.vcpenc_session_upload_async
        self._ck(fn(self.h, _ptr(frames), nframes, self.err, 512))
        self.nframes = nframes

    def upload_device(self, dptr: int, nframes: int) -> float:
        """Raw frames already in HBM (device pointer): runs K1 only; returns its CUDA-event ms."""
        ms = C.c_float(0)
        self._ck(self.L.vcpenc_session_upload_device(self.h, dptr, nframes, C.byref(ms), self.err, 512))
        self.nframes = nframes
        return ms.value

    def launch_count(self) -> int:
        return int(self.L.vcpenc_session_launch_count(self.h))

    def encode(self) -> float:
        ms = C.c_float(0)
        self._ck(self.L.vcpenc_session_encode(self.h, C.byref(ms), self.err, 512))
        return ms.value

    def download(self, want_recon=False, out=None):
        n = self.nframes
        if out is None:
            out = np.empty(n * self.fb + (1 << 20), np.uint8)
        out_len = C.c_size_t(0)
        info = (FrameInfo * n)()
        recon = np.empty((n, self.fb), np.uint8) if want_recon else None
        self._ck(self.L.vcpenc_session_download(self.h, _ptr(out), out.size, C.byref(out_len),
                                                C.cast(info, C.c_void_p), _ptr(recon), self.err, 512))
        return {"stream": out[: out_len.value], "info": [(i.offset, i.size, i.is_idr, i.qp) for i in info],
                "recon": recon}

    def profile(self, enable=True):
        self.L.vcpenc_session_profile(self.h, 1 if enable else 0)

    def kernel_stats(self):
        st = (KernelStat * len(K_NAMES))()
        self.L.vcpenc_session_kernel_stats(self.h, C.cast(st, C.c_void_p))
        return {K_NAMES[i]: {"ms": st[i].ms, "launches": int(st[i].launches)} for i in range(len(K_NAMES))}

    def debug_mbs(self):
        n = self.nframes
        d = {"mv_prepass": np.zeros((n, self.mbh, self.mbw, 2), np.int16),
             "mv_final": np.zeros((n, self.mbh, self.mbw, 2), np.int16),
             "mb_type": np.zeros((n, self.mbh, self.mbw), np.uint8),
             "cbp": np.zeros((n, self.mbh, self.mbw), np.uint8)}
        rc = self.L.vcpenc_session_debug_mbs(self.h, d["mv_prepass"].ctypes.data, d["mv_final"].ctypes.data,
                                             d["mb_type"].ctypes.data, d["cbp"].ctypes.data)
        if rc:
            raise VcpencError(rc, "debug taps unavailable (params.debug=1 and encode first)")
        d["mv_prepass"] *= 4  # full-pel -> quarter-pel units, as the oracle reports them
        return d

    def close(self):
        if self.h:
            self.L.vcpenc_session_destroy(self.h)
            self.h = C.c_void_p(None)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def set_thread_device(device: int):
    """Bind transcode() calls of the calling thread to a GPU (one consumer worker per device)."""
    rc = lib().vcpenc_set_thread_device(int(device))
    if rc:
        raise VcpencError(rc, "device %d not available" % device)


def ffmpeg_libdir():
    """Where a loadable libavformat/libavcodec lives in this image (the opencv-bundled LGPL build);
    exported as VCPENC_FFMPEG_LIBDIR for the container front end unless the operator set one."""
    from . import arbiter
    d = arbiter._find_libdir()
    if d and "VCPENC_FFMPEG_LIBDIR" not in os.environ:
        os.environ["VCPENC_FFMPEG_LIBDIR"] = d
    return os.environ.get("VCPENC_FFMPEG_LIBDIR")


def probe_input(path, max_frames=0):
    """Geometry (and optionally decoded pictures) of a container input, as transcode() will see it."""
    L = lib()
    ffmpeg_libdir()
    L.vcpenc_probe_input.argtypes = [C.c_char_p] + [C.POINTER(C.c_int)] * 5 + [C.c_void_p, C.c_size_t, C.c_int,
                                                                             C.POINTER(C.c_int), C.c_char_p, C.c_size_t]
    w, h, fn, fd, fmt, n = (C.c_int(0) for _ in range(6))
    err = C.create_string_buffer(512)
    rc = L.vcpenc_probe_input(os.fsencode(path), C.byref(w), C.byref(h), C.byref(fn), C.byref(fd), C.byref(fmt),
                              None, 0, 0, C.byref(n), err, 512)
    if rc:
        raise VcpencError(rc, err.value.decode(errors="replace"))
    info = {"width": w.value, "height": h.value, "fps": (fn.value, fd.value), "fmt": fmt.value}
    if max_frames > 0:
        p = Params()
        p.in_fmt, p.width, p.height = fmt.value, w.value, h.value
        fb = in_frame_bytes(p)
        buf = np.empty(max_frames * fb, np.uint8)
        rc = L.vcpenc_probe_input(os.fsencode(path), None, None, None, None, None, buf.ctypes.data, buf.size, max_frames,
                                  C.byref(n), err, 512)
        if rc:
            raise VcpencError(rc, err.value.decode(errors="replace"))
        info["frames"] = buf[: n.value * fb].reshape(n.value, fb)
    return info


def probe_audio(path, audio_bitrate=0):
    """The audio side of a container input as transcode() handles it: dict(sample_rate, channels, copied, priming,
    asc, frames=[bytes...]) -- raw AAC access units (stream copy of AAC-LC, else decode + libavcodec aac)."""
    L = lib()
    ffmpeg_libdir()
    ip = C.POINTER(C.c_int)
    L.vcpenc_probe_audio.argtypes = [C.c_char_p, C.c_int, ip, ip, ip, ip, C.c_void_p, C.c_int, ip, C.c_void_p, C.c_size_t,
                                     C.POINTER(C.c_size_t), C.c_void_p, C.c_int, ip, C.c_char_p, C.c_size_t]
    rate, ch, copied, prim, alen, nfr = (C.c_int(0) for _ in range(6))
    dlen = C.c_size_t(0)
    err = C.create_string_buffer(512)
    asc = np.zeros(64, np.uint8)
    cap = 64 << 20
    data = np.empty(cap, np.uint8)
    sizes = np.empty(1 << 20, np.uint32)
    rc = L.vcpenc_probe_audio(os.fsencode(path), int(audio_bitrate), C.byref(rate), C.byref(ch), C.byref(copied), C.byref(prim),
                              asc.ctypes.data, 64, C.byref(alen), data.ctypes.data, cap, C.byref(dlen), sizes.ctypes.data, sizes.size,
                              C.byref(nfr), err, 512)
    if rc:
        raise VcpencError(rc, err.value.decode(errors="replace"))
    frames, o = [], 0
    for k in range(nfr.value):
        frames.append(data[o:o + int(sizes[k])].tobytes())
        o += int(sizes[k])
    return {"sample_rate": rate.value, "channels": ch.value, "copied": bool(copied.value), "priming": prim.value,
            "asc": asc[:alen.value].tobytes(), "frames": frames}


def thread_release():
    """Free the session / pinned buffer transcode() keeps for the calling thread."""
    lib().vcpenc_thread_release()


def parse_args(tokens):
    """strings.Fields(ffmpeg_args) -> Params (raises VcpencError, e.g. NOTENCODE for `-c copy`)."""
    L = lib()
    arr = (C.c_char_p * len(tokens))(*[t.encode() for t in tokens])
    p = Params()
    err = C.create_string_buffer(512)
    rc = L.vcpenc_parse_args(len(tokens), arr, C.byref(p), err, 512)
    if rc:
        raise VcpencError(rc, err.value.decode(errors="replace"))
    return p


def transcode(input_path, output_path, ffmpeg_args: str, timeout_ms=60 * 60 * 1000, cancel=None):
    """Mirror of runFFmpegWithTimeout (cmd/consumer.go:370-394): ffmpeg_args is split on
    whitespace only (strings.Fields), `-y` overwrite semantics, error classes for timeout
    and cancellation.  `cancel` is an optional ctypes.c_int polled by the library."""
    L = lib()
    ffmpeg_libdir()
    toks = ffmpeg_args.split()
    arr = (C.c_char_p * max(1, len(toks)))(*[t.encode() for t in toks])
    err = C.create_string_buffer(1024)
    rc = L.vcpenc_transcode(os.fsencode(input_path), os.fsencode(output_path), len(toks), arr, int(timeout_ms),
                            C.byref(cancel) if cancel is not None else None, err, 1024)
    if rc:
        raise VcpencError(rc, err.value.decode(errors="replace"))


def verify(path):
    """Mirror of verifyOutputFile (cmd/consumer.go:396-419)."""
    L = lib()
    err = C.create_string_buffer(512)
    rc = L.vcpenc_verify(os.fsencode(path), err, 512)
    if rc:
        raise VcpencError(rc, err.value.decode(errors="replace"))


def mux_mp4(params: Params, stream: np.ndarray, info, path, audio=None):
    """audio: dict as returned by probe_audio() -> an AAC track beside the video."""
    L = lib()
    if audio is not None:
        return _mux_mp4_audio(L, params, stream, info, path, audio)
    n = len(info)
    fi = (FrameInfo * n)()
    for k, (off, size, idr, qp) in enumerate(info):
        fi[k].offset, fi[k].size, fi[k].is_idr, fi[k].qp = off, size, idr, qp
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    err = C.create_string_buffer(512)
    rc = L.vcpenc_mux_mp4(C.byref(params), stream.ctypes.data, stream.size, C.cast(fi, C.c_void_p), n,
                          os.fsencode(path), err, 512)
    if rc:
        raise VcpencError(rc, err.value.decode(errors="replace"))


def _mux_mp4_audio(L, params, stream, info, path, audio):
    n = len(info)
    fi = (FrameInfo * n)()
    for k, (off, size, idr, qp) in enumerate(info):
        fi[k].offset, fi[k].size, fi[k].is_idr, fi[k].qp = off, size, idr, qp
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    aac = np.frombuffer(b"".join(audio["frames"]), np.uint8)
    sizes = np.array([len(f) for f in audio["frames"]], np.uint32)
    asc = np.frombuffer(audio["asc"], np.uint8)
    L.vcpenc_mux_mp4_audio.argtypes = [C.POINTER(Params), C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_size_t]
    err = C.create_string_buffer(512)
    rc = L.vcpenc_mux_mp4_audio(C.byref(params), stream.ctypes.data, stream.size, C.cast(fi, C.c_void_p), n, aac.ctypes.data,
                                sizes.ctypes.data, len(sizes), audio["sample_rate"], audio["channels"], audio["priming"],
                                asc.ctypes.data, asc.size, os.fsencode(path), err, 512)
    if rc:
        raise VcpencError(rc, err.value.decode(errors="replace"))
