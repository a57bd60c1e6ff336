"""Arbiter: the FFmpeg (libavcodec/libavformat) decoder, reached header-less over ctypes.

This is VERIFICATION infrastructure, not part of the encode path.  The north-star's
correctness criterion is "the reference FFmpeg decoder must reproduce the encoder's own
reconstructed frames exactly"; the only FFmpeg in this image is the LGPL build bundled
inside opencv-python-headless, which has the native `h264`
decoder and the `mov,mp4` demuxer but no headers.  So only allocation/accessor APIs plus
the stable heads of AVFrame / AVPacket / AVFormatContext / AVStream are touched.

It plays the role of the two external binaries the reference shells out to:
  * `ffmpeg` decode side  -> decode_annexb() / decode_file()      (bit-exactness, PSNR)
  * `ffprobe -select_streams v:0 -show_entries stream=codec_type`
                           -> probe_has_video()   (/root/reference/cmd/consumer.go:409-418)
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import numpy as np

_LIBS = None


def _find_libdir():
    import importlib.util
    spec = importlib.util.find_spec("cv2")
    if spec is None or not spec.submodule_search_locations:
        return None
    site = os.path.dirname(list(spec.submodule_search_locations)[0])
    d = os.path.join(site, "opencv_python_headless.libs")
    return d if os.path.isdir(d) else None


def _load():
    global _LIBS
    if _LIBS is not None:
        return _LIBS
    d = _find_libdir()
    if d is None:
        raise RuntimeError("bundled FFmpeg libraries (opencv_python_headless.libs) not found")

    # The bundled libraries depend on each other by hashed soname without an rpath that
    # ctypes honours: preload the whole directory, retrying until dependencies resolve.
    pending = sorted(glob.glob(os.path.join(d, "*.so*")))
    for _ in range(6):
        nxt = []
        for path in pending:
            try:
                C.CDLL(path, mode=C.RTLD_GLOBAL)
            except OSError:
                nxt.append(path)
        if not nxt or len(nxt) == len(pending):
            break
        pending = nxt

    def one(stem):
        m = sorted(glob.glob(os.path.join(d, stem + "-*.so*")))
        if not m:
            raise RuntimeError("missing " + stem)
        return C.CDLL(m[0], mode=C.RTLD_GLOBAL)

    avutil = one("libavutil")
    one("libswresample")
    avcodec = one("libavcodec")
    avformat = one("libavformat")
    vp = C.c_void_p
    avcodec.avcodec_find_decoder_by_name.restype = vp
    avcodec.avcodec_find_decoder_by_name.argtypes = [C.c_char_p]
    avcodec.avcodec_find_decoder.restype = vp
    avcodec.avcodec_find_decoder.argtypes = [C.c_int]
    avcodec.avcodec_alloc_context3.restype = vp
    avcodec.avcodec_alloc_context3.argtypes = [vp]
    avcodec.avcodec_open2.argtypes = [vp, vp, vp]
    avcodec.avcodec_free_context.argtypes = [vp]
    avcodec.av_packet_alloc.restype = vp
    avcodec.av_new_packet.argtypes = [vp, C.c_int]
    avcodec.av_packet_unref.argtypes = [vp]
    avcodec.av_packet_free.argtypes = [vp]
    avcodec.avcodec_send_packet.argtypes = [vp, vp]
    avcodec.avcodec_receive_frame.argtypes = [vp, vp]
    avcodec.avcodec_parameters_to_context.argtypes = [vp, vp]
    avutil.av_frame_alloc.restype = vp
    avutil.av_frame_unref.argtypes = [vp]
    avutil.av_frame_free.argtypes = [vp]
    avutil.av_opt_set_int.argtypes = [vp, C.c_char_p, C.c_int64, C.c_int]
    avutil.av_opt_set.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int]
    avformat.avformat_open_input.argtypes = [C.POINTER(vp), C.c_char_p, vp, vp]
    avformat.avformat_find_stream_info.argtypes = [vp, vp]
    avformat.avformat_close_input.argtypes = [C.POINTER(vp)]
    avformat.av_find_best_stream.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int]
    avformat.av_read_frame.argtypes = [vp, vp]
    _LIBS = (avutil, avcodec, avformat)
    return _LIBS


def available() -> bool:
    try:
        _load()
        return True
    except Exception:
        return False


# ---- struct heads (FFmpeg 6/7/8 ABI) -------------------
class _AVFrameHead(C.Structure):
    _fields_ = [("data", C.c_void_p * 8), ("linesize", C.c_int * 8),
                ("extended_data", C.c_void_p), ("width", C.c_int), ("height", C.c_int),
                ("nb_samples", C.c_int), ("format", C.c_int)]


class _AVPacketHead(C.Structure):
    _fields_ = [("buf", C.c_void_p), ("pts", C.c_int64), ("dts", C.c_int64),
                ("data", C.c_void_p), ("size", C.c_int), ("stream_index", C.c_int)]


AV_PIX_FMT_YUV420P = 0
AV_PIX_FMT_YUVJ420P = 12
AVERROR_EAGAIN = -11
AVMEDIA_TYPE_VIDEO = 0


def _frame_to_planes(fr: _AVFrameHead):
    if fr.format not in (AV_PIX_FMT_YUV420P, AV_PIX_FMT_YUVJ420P):
        raise RuntimeError(f"unexpected pix_fmt {fr.format}")
    w, h = fr.width, fr.height
    out = []
    for i, (pw, ph) in enumerate(((w, h), ((w + 1) // 2, (h + 1) // 2), ((w + 1) // 2, (h + 1) // 2))):
        ls = fr.linesize[i]
        buf = (C.c_uint8 * (ls * ph)).from_address(fr.data[i])
        a = np.frombuffer(buf, dtype=np.uint8).reshape(ph, ls)[:, :pw].copy()
        out.append(a)
    return out


class _Decoder:
    def __init__(self, name=b"h264", threads=0, codecpar=None):
        avutil, avcodec, _ = _load()
        self.avutil, self.avcodec = avutil, avcodec
        if name is None and codecpar:
            codec = avcodec.avcodec_find_decoder(C.c_int.from_address(codecpar + 4).value)   # AVCodecParameters.codec_id
        else:
            codec = avcodec.avcodec_find_decoder_by_name(name)
        if not codec:
            raise RuntimeError("decoder not found: %r" % name)
        self.ctx = avcodec.avcodec_alloc_context3(codec)
        if codecpar:
            if avcodec.avcodec_parameters_to_context(self.ctx, codecpar) < 0:
                raise RuntimeError("avcodec_parameters_to_context failed")
        avutil.av_opt_set_int(self.ctx, b"threads", threads, 0)
        # bit-exact, no error concealment surprises
        avutil.av_opt_set(self.ctx, b"flags", b"+bitexact", 0)
        r = avcodec.avcodec_open2(self.ctx, codec, None)
        if r < 0:
            raise RuntimeError("avcodec_open2 failed %d" % r)
        self.pkt = avcodec.av_packet_alloc()
        self.frame = avutil.av_frame_alloc()
        self.frames = []

    def _drain(self):
        while True:
            r = self.avcodec.avcodec_receive_frame(self.ctx, self.frame)
            if r < 0:
                return r
            fr = _AVFrameHead.from_address(self.frame)
            self.frames.append(_frame_to_planes(fr))
            self.avutil.av_frame_unref(self.frame)

    def send(self, data: bytes):
        n = len(data)
        r = self.avcodec.av_new_packet(self.pkt, n)
        if r < 0:
            raise RuntimeError("av_new_packet failed")
        ph = _AVPacketHead.from_address(self.pkt)
        C.memmove(ph.data, data, n)
        r = self.avcodec.avcodec_send_packet(self.ctx, self.pkt)
        self.avcodec.av_packet_unref(self.pkt)
        if r < 0 and r != AVERROR_EAGAIN:
            raise RuntimeError("avcodec_send_packet failed %d" % r)
        self._drain()

    def send_pkt(self, pkt):
        r = self.avcodec.avcodec_send_packet(self.ctx, pkt)
        if r < 0 and r != AVERROR_EAGAIN:
            raise RuntimeError("avcodec_send_packet failed %d" % r)
        self._drain()

    def flush(self):
        self.avcodec.avcodec_send_packet(self.ctx, None)
        self._drain()

    def close(self):
        p = C.c_void_p(self.frame)
        self.avutil.av_frame_free(C.byref(p))
        p = C.c_void_p(self.pkt)
        self.avcodec.av_packet_free(C.byref(p))
        p = C.c_void_p(self.ctx)
        self.avcodec.avcodec_free_context(C.byref(p))


def split_annexb(data: bytes):
    """Split an Annex-B byte stream into NAL units (without start codes)."""
    nals = []
    i = 0
    n = len(data)
    starts = []
    while True:
        j = data.find(b"\x00\x00\x01", i)
        if j < 0:
            break
        starts.append(j + 3)
        i = j + 3
    for k, s in enumerate(starts):
        e = starts[k + 1] - 3 if k + 1 < len(starts) else n
        # strip trailing zero bytes belonging to the next 4-byte start code
        while e > s and data[e - 1] == 0:
            e -= 1
        nals.append(data[s:e])
    return nals


def group_access_units(nals):
    """Group NAL units into access units (one coded picture each), Annex-B bytes out."""
    aus = []
    cur = []
    have_vcl = False
    for nal in nals:
        t = nal[0] & 0x1F
        is_vcl = t in (1, 5)
        new_pic = False
        if is_vcl:
            first_mb_is_zero = (nal[1] & 0x80) != 0  # ue(0) == '1'
            if have_vcl and first_mb_is_zero:
                new_pic = True
        elif have_vcl:
            new_pic = True
        if new_pic:
            aus.append(b"".join(b"\x00\x00\x00\x01" + x for x in cur))
            cur = []
            have_vcl = False
        cur.append(nal)
        have_vcl = have_vcl or is_vcl
    if cur:
        aus.append(b"".join(b"\x00\x00\x00\x01" + x for x in cur))
    return aus


def decode_annexb(data: bytes, codec=b"h264", threads=0):
    """Decode an Annex-B elementary stream; returns a list of [Y,U,V] uint8 planes per frame."""
    dec = _Decoder(codec, threads)
    try:
        for au in group_access_units(split_annexb(data)):
            dec.send(au)
        dec.flush()
        return dec.frames
    finally:
        dec.close()


def group_access_units_hevc(nals):
    """HEVC: a new access unit starts at a VCL NAL with first_slice_segment_in_pic_flag, or at a parameter set /
    AUD after VCL data."""
    aus, cur, have_vcl = [], [], False
    for nal in nals:
        t = (nal[0] >> 1) & 63
        is_vcl = t < 32
        new_pic = (is_vcl and have_vcl and (nal[2] & 0x80)) or (not is_vcl and have_vcl)
        if new_pic:
            aus.append(b"".join(b"\x00\x00\x00\x01" + x for x in cur))
            cur, have_vcl = [], False
        cur.append(nal)
        have_vcl = have_vcl or is_vcl
    if cur:
        aus.append(b"".join(b"\x00\x00\x00\x01" + x for x in cur))
    return aus


def decode_annexb_hevc(data: bytes, threads=0):
    """Decode an HEVC Annex-B elementary stream with the FFmpeg `hevc` decoder."""
    dec = _Decoder(b"hevc", threads)
    try:
        for au in group_access_units_hevc(split_annexb(data)):
            dec.send(au)
        dec.flush()
        return dec.frames
    finally:
        dec.close()


def _open_input(path):
    _, _, avformat = _load()
    fmt = C.c_void_p(None)
    r = avformat.avformat_open_input(C.byref(fmt), os.fsencode(path), None, None)
    if r < 0:
        raise RuntimeError("avformat_open_input failed %d" % r)
    r = avformat.avformat_find_stream_info(fmt, None)
    if r < 0:
        avformat.avformat_close_input(C.byref(fmt))
        raise RuntimeError("avformat_find_stream_info failed %d" % r)
    return fmt


def probe_has_video(path) -> bool:
    """ffprobe-equivalent of the reference's --verify (cmd/consumer.go:396-419)."""
    try:
        if os.path.getsize(path) == 0:
            return False
        _, _, avformat = _load()
        fmt = _open_input(path)
    except Exception:
        return False
    try:
        return avformat.av_find_best_stream(fmt, AVMEDIA_TYPE_VIDEO, -1, -1, None, 0) >= 0
    finally:
        avformat.avformat_close_input(C.byref(fmt))


def decode_file(path, threads=0):
    """Demux + decode the best video stream of a container file (mp4/mkv/...)."""
    _, avcodec, avformat = _load()
    fmt = _open_input(path)
    try:
        si = avformat.av_find_best_stream(fmt, AVMEDIA_TYPE_VIDEO, -1, -1, None, 0)
        if si < 0:
            raise RuntimeError("no video stream")
        # AVFormatContext: av_class,iformat,oformat,priv_data,pb (5 ptr) ctx_flags,nb_streams (2 int) streams
        streams = C.c_void_p.from_address(fmt.value + 48).value
        st = C.c_void_p.from_address(streams + 8 * si).value
        # AVStream: av_class (ptr) index,id (2 int) codecpar (ptr)
        codecpar = C.c_void_p.from_address(st + 16).value
        codec_id = C.c_int.from_address(codecpar + 4).value
        codec = avcodec.avcodec_find_decoder(codec_id)
        if not codec:
            raise RuntimeError("no decoder for codec id %d" % codec_id)
        dec = _Decoder.__new__(_Decoder)
        avutil = _load()[0]
        dec.avutil, dec.avcodec = avutil, avcodec
        dec.ctx = avcodec.avcodec_alloc_context3(codec)
        avcodec.avcodec_parameters_to_context(dec.ctx, codecpar)
        avutil.av_opt_set_int(dec.ctx, b"threads", threads, 0)
        if avcodec.avcodec_open2(dec.ctx, codec, None) < 0:
            raise RuntimeError("avcodec_open2 failed")
        dec.pkt = avcodec.av_packet_alloc()
        dec.frame = avutil.av_frame_alloc()
        dec.frames = []
        try:
            while avformat.av_read_frame(fmt, dec.pkt) >= 0:
                ph = _AVPacketHead.from_address(dec.pkt)
                if ph.stream_index == si:
                    dec.send_pkt(dec.pkt)
                avcodec.av_packet_unref(dec.pkt)
            dec.flush()
            return dec.frames
        finally:
            dec.close()
    finally:
        avformat.avformat_close_input(C.byref(fmt))


def _audio_frame_to_float(fr: _AVFrameHead, channels: int) -> np.ndarray:
    """decoded audio AVFrame -> float32 [nb_samples, channels] (fltp, flt, s16, s16p)."""
    n = fr.nb_samples
    ext = C.cast(fr.extended_data, C.POINTER(C.c_void_p))
    if fr.format == 8:      # fltp
        return np.stack([np.frombuffer((C.c_float * n).from_address(ext[c]), np.float32).copy() for c in range(channels)], axis=1)
    if fr.format == 3:      # flt
        return np.frombuffer((C.c_float * (n * channels)).from_address(ext[0]), np.float32).copy().reshape(n, channels)
    if fr.format == 1:      # s16
        return np.frombuffer((C.c_int16 * (n * channels)).from_address(ext[0]), np.int16).astype(np.float32).reshape(n, channels) / 32768.0
    if fr.format == 6:      # s16p
        return np.stack([np.frombuffer((C.c_int16 * n).from_address(ext[c]), np.int16).astype(np.float32) / 32768.0 for c in range(channels)], axis=1)
    raise RuntimeError("unexpected sample format %d" % fr.format)


def decode_adts(data: bytes, channels: int) -> np.ndarray:
    """ADTS AAC stream -> float32 [samples, channels] through libavcodec's aac decoder."""
    avutil, avcodec, _ = _load()
    dec = _Decoder(b"aac")
    out = []
    o = 0
    pkts = []
    while o + 7 <= len(data):
        n = ((data[o + 3] & 3) << 11) | (data[o + 4] << 3) | (data[o + 5] >> 5)
        if n < 7 or o + n > len(data):
            break
        pkts.append(data[o:o + n])
        o += n
    frame = dec.frame
    def drain():
        while avcodec.avcodec_receive_frame(dec.ctx, frame) >= 0:
            out.append(_audio_frame_to_float(_AVFrameHead.from_address(frame), channels))
            avutil.av_frame_unref(frame)
    for p in pkts:
        avcodec.av_new_packet(dec.pkt, len(p))
        ph = _AVPacketHead.from_address(dec.pkt)
        C.memmove(ph.data, p, len(p))
        avcodec.avcodec_send_packet(dec.ctx, dec.pkt)
        avcodec.av_packet_unref(dec.pkt)
        drain()
    avcodec.avcodec_send_packet(dec.ctx, None)
    drain()
    dec.close()
    return np.concatenate(out, axis=0) if out else np.zeros((0, channels), np.float32)


def decode_audio_file(path):
    """(first audio stream of a container file) -> (codec_id, float32 [samples, channels]) via libavformat + libavcodec;
    None if the file has no audio stream."""
    avutil, avcodec, avformat = _load()
    ctx = _open_input(path)
    try:
        ai = avformat.av_find_best_stream(ctx, 1, -1, -1, None, 0)
        if ai < 0:
            return None
        streams = C.cast(C.c_void_p.from_address(ctx.value + 48).value, C.POINTER(C.c_void_p))
        st = streams[ai]
        par = C.c_void_p.from_address(st + 16).value
        codec_id = C.c_int.from_address(par + 4).value
        dec = _Decoder(None, codecpar=par)
        avutil.av_opt_get_chlayout.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p]
        lay = (C.c_int * 6)()
        avutil.av_opt_get_chlayout(dec.ctx, b"ch_layout", 0, lay)
        channels = max(1, lay[1])
        out = []
        frame = dec.frame
        def drain():
            while avcodec.avcodec_receive_frame(dec.ctx, frame) >= 0:
                out.append(_audio_frame_to_float(_AVFrameHead.from_address(frame), channels))
                avutil.av_frame_unref(frame)
        pkt = dec.pkt
        while avformat.av_read_frame(ctx, pkt) >= 0:
            ph = _AVPacketHead.from_address(pkt)
            if ph.stream_index == ai:
                avcodec.avcodec_send_packet(dec.ctx, pkt)
                drain()
            avcodec.av_packet_unref(pkt)
        avcodec.avcodec_send_packet(dec.ctx, None)
        drain()
        dec.close()
        return codec_id, (np.concatenate(out, axis=0) if out else np.zeros((0, channels), np.float32))
    finally:
        avformat.avformat_close_input(C.byref(ctx))


def psnr(a: np.ndarray, b: np.ndarray) -> float:
    d = a.astype(np.int64) - b.astype(np.int64)
    mse = float((d * d).mean())
    return 99.0 if mse == 0 else 10.0 * np.log10(255.0 * 255.0 / mse)
