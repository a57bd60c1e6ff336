"""Test inputs with an audio track, written without FFmpeg: AVI (RIFF) holding raw I420 pictures and 16-bit
PCM.  libavformat's `avi` demuxer and libavcodec's `rawvideo` / `pcm_s16le` decoders read it, which makes it a
foreign container + foreign codecs + audio for the front end."""
from __future__ import annotations

import struct

import numpy as np


def tone(rate: int, seconds: float, channels: int = 2, freqs=(300.0, 500.0)) -> np.ndarray:
    """int16 [n, channels]: one linear chirp per channel (start frequency per channel, two octaves up over the clip),
    about -9 dBFS: aperiodic, so a cross-correlation against the decoded output has a single peak."""
    n = int(rate * seconds)
    t = np.arange(n) / rate
    cols = []
    for c in range(channels):
        f0 = freqs[c % len(freqs)]
        phase = 2 * np.pi * (f0 * t + 1.5 * f0 * t * t / max(seconds, 1e-9))
        cols.append(0.35 * np.sin(phase))
    return np.round(np.stack(cols, axis=1) * 32767).astype(np.int16)


def _chunk(tag: bytes, data: bytes) -> bytes:
    return tag + struct.pack("<I", len(data)) + data + (b"\0" if len(data) & 1 else b"")


def _list(kind: bytes, data: bytes) -> bytes:
    return b"LIST" + struct.pack("<I", len(data) + 4) + kind + data


def write_avi(path, frames: np.ndarray, w: int, h: int, fps: int, pcm: np.ndarray | None = None, rate: int = 48000):
    """frames: uint8 [n, w*h*3/2] yuv420p (I420); pcm: int16 [samples, channels] or None."""
    n = frames.shape[0]
    fb = w * h * 3 // 2
    assert frames.shape[1] == fb
    nch = pcm.shape[1] if pcm is not None else 0
    streams = 1 + (1 if pcm is not None else 0)
    avih = struct.pack("<IIIIIIIIII4I", 1000000 // fps, fb * fps, 0, 0x10, n, 0, streams, fb, w, h, 0, 0, 0, 0)
    strh_v = struct.pack("<4s4sIHHIIIIIIIIhhhh", b"vids", b"I420", 0, 0, 0, 0, 1, fps, 0, n, fb, 0xFFFFFFFF, 0, 0, 0, w, h)
    strf_v = struct.pack("<IiiHH4sIiiII", 40, w, h, 1, 12, b"I420", fb, 0, 0, 0, 0)
    hdrl = _chunk(b"avih", avih) + _list(b"strl", _chunk(b"strh", strh_v) + _chunk(b"strf", strf_v))
    if pcm is not None:
        ba = 2 * nch
        strh_a = struct.pack("<4s4sIHHIIIIIIIIhhhh", b"auds", b"\0\0\0\0", 0, 0, 0, 0, ba, rate * ba, 0, pcm.shape[0], rate * ba, 0xFFFFFFFF, ba, 0, 0, 0, 0)
        strf_a = struct.pack("<HHIIHH", 1, nch, rate, rate * ba, ba, 16)
        hdrl += _list(b"strl", _chunk(b"strh", strh_a) + _chunk(b"strf", strf_a))
    movi = b""
    idx = b""
    off = 4                     # offsets in idx1 are relative to the 'movi' fourcc
    per = (pcm.shape[0] + n - 1) // n if pcm is not None else 0
    for i in range(n):
        c = _chunk(b"00dc", frames[i].tobytes())
        idx += struct.pack("<4sIII", b"00dc", 0x10, off, fb)
        movi += c
        off += len(c)
        if pcm is not None:
            blk = pcm[i * per:(i + 1) * per].tobytes()
            if blk:
                c = _chunk(b"01wb", blk)
                idx += struct.pack("<4sIII", b"01wb", 0x10, off, len(blk))
                movi += c
                off += len(c)
    body = b"AVI " + _list(b"hdrl", hdrl) + _list(b"movi", movi) + _chunk(b"idx1", idx)
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(body)) + body)


def adts_wrap(frames, rate: int, channels: int) -> bytes:
    """Raw AAC-LC access units -> ADTS stream (what libavcodec's `aac` decoder takes without extradata)."""
    rates = [96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000, 12000, 11025, 8000, 7350]
    fi = rates.index(rate)
    cc = 7 if channels == 8 else channels
    out = bytearray()
    for fr in frames:
        n = len(fr) + 7
        out += bytes([0xFF, 0xF1, (1 << 6) | (fi << 2) | (cc >> 2), ((cc & 3) << 6) | (n >> 11), (n >> 3) & 0xFF, ((n & 7) << 5) | 0x1F, 0xFC])
        out += fr
    return bytes(out)


def best_lag_snr(ref: np.ndarray, got: np.ndarray, max_lag: int = 4096):
    """Align `got` to `ref` (1-D float arrays) by cross-correlation over +-max_lag samples; returns (lag, snr_db)."""
    n = min(len(ref), len(got)) - 2 * max_lag
    a = ref[max_lag:max_lag + n].astype(np.float64)
    best = (0, -1e9)
    # coarse search on a decimated signal, then exact around the peak
    cands = range(-max_lag, max_lag + 1, 8)
    def corr(lag):
        b = got[max_lag + lag:max_lag + lag + n].astype(np.float64)
        return float(np.dot(a, b))
    c0 = max(cands, key=corr)
    lag = max(range(c0 - 8, c0 + 9), key=corr)
    b = got[max_lag + lag:max_lag + lag + n].astype(np.float64)
    g = np.dot(a, b) / max(np.dot(b, b), 1e-12)
    err = a - g * b
    snr = 10 * np.log10(max(np.dot(a, a), 1e-12) / max(np.dot(err, err), 1e-12))
    return lag, snr
