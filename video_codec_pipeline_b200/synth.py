"""Deterministic synthetic yuv420p clips.

libavfilter is absent from the image, so FFmpeg's `testsrc2` cannot be generated; this is
the stand-in the survey describes: colour bars + a diagonal gradient scrolling 4 px/frame +
a 256x256 uniform-noise patch translating (+3,+2) px/frame + a frame counter, 8-bit
yuv420p, generator numpy default_rng(seed).  The seed is reported in the bench JSON.
"""
from __future__ import annotations

import numpy as np

_DIGITS = {
    "0": ("111", "101", "101", "101", "111"), "1": ("010", "110", "010", "010", "111"),
    "2": ("111", "001", "111", "100", "111"), "3": ("111", "001", "111", "001", "111"),
    "4": ("101", "101", "111", "001", "001"), "5": ("111", "100", "111", "001", "111"),
    "6": ("111", "100", "111", "101", "111"), "7": ("111", "001", "010", "010", "010"),
    "8": ("111", "101", "111", "101", "111"), "9": ("111", "101", "111", "001", "111"),
}


def frame_bytes(w: int, h: int) -> int:
    return w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2)


def _bars(w, h):
    # 8 vertical bars, BT.709 limited-range-ish YUV triples
    cols = np.array([[235, 128, 128], [210, 16, 146], [170, 166, 16], [145, 54, 34],
                     [106, 202, 222], [81, 90, 240], [41, 240, 110], [16, 128, 128]], np.uint8)
    idx = (np.arange(w) * 8 // w).astype(np.int64)
    return cols[idx]  # [w,3]


def make_clip(w: int, h: int, nframes: int, seed: int, start: int = 0) -> np.ndarray:
    """Returns uint8 [nframes, frame_bytes(w,h)] planar yuv420p."""
    rng = np.random.default_rng(seed)
    ps = min(256, (min(w, h) // 2) & ~1)
    noise = rng.integers(0, 256, size=(ps, ps), dtype=np.uint8)
    noise_u = rng.integers(96, 160, size=(ps // 2, ps // 2), dtype=np.uint8)
    noise_v = rng.integers(96, 160, size=(ps // 2, ps // 2), dtype=np.uint8)
    grain = rng.integers(-2, 3, size=(h, w), dtype=np.int16)
    bars = _bars(w, h)
    cw, ch = (w + 1) // 2, (h + 1) // 2
    out = np.empty((nframes, frame_bytes(w, h)), np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    for i in range(nframes):
        n = start + i
        Y = np.empty((h, w), np.int16)
        U = np.empty((ch, cw), np.int16)
        V = np.empty((ch, cw), np.int16)
        top = h // 3
        # colour bars (static) in the top third
        Y[:top] = bars[:, 0][None, :]
        U[: top // 2] = bars[::2, 1][None, :cw]
        V[: top // 2] = bars[::2, 2][None, :cw]
        # scrolling diagonal gradient below
        g = ((xx[top:] + yy[top:] + 4 * n) % 512)
        g = np.where(g < 256, g, 511 - g)
        Y[top:] = 16 + (g * 219 // 255)
        U[top // 2:] = 128 + ((((xx[top:: 2, ::2][: ch - top // 2, :cw] + 2 * n) % 256) - 128) // 4)
        V[top // 2:] = 128 - ((((yy[top:: 2, ::2][: ch - top // 2, :cw] + 2 * n) % 256) - 128) // 4)
        Y += grain  # fixed-pattern grain so flat areas are not trivially flat
        # translating noise patch
        px = (w // 8 + 3 * n) % max(1, w - ps)
        py = (h // 8 + 2 * n) % max(1, h - ps)
        px &= ~1
        py &= ~1
        Y[py:py + ps, px:px + ps] = noise
        U[py // 2:py // 2 + ps // 2, px // 2:px // 2 + ps // 2] = noise_u
        V[py // 2:py // 2 + ps // 2, px // 2:px // 2 + ps // 2] = noise_v
        # frame counter, 8 digits, 3x5 font scaled
        sc = max(2, h // 135)
        s = "%08d" % n
        x0, y0 = w // 16, h - 8 * sc
        for k, c in enumerate(s):
            for r, row in enumerate(_DIGITS[c]):
                for q, bit in enumerate(row):
                    if bit == "1":
                        ys, xs = y0 + r * sc, x0 + (k * 4 + q) * sc
                        Y[ys:ys + sc, xs:xs + sc] = 235
        f = out[i]
        f[: w * h] = np.clip(Y, 0, 255).astype(np.uint8).ravel()
        f[w * h: w * h + cw * ch] = np.clip(U, 0, 255).astype(np.uint8).ravel()
        f[w * h + cw * ch:] = np.clip(V, 0, 255).astype(np.uint8).ravel()
    return out


def split_planes(frame: np.ndarray, w: int, h: int):
    cw, ch = (w + 1) // 2, (h + 1) // 2
    y = frame[: w * h].reshape(h, w)
    u = frame[w * h: w * h + cw * ch].reshape(ch, cw)
    v = frame[w * h + cw * ch:].reshape(ch, cw)
    return y, u, v


def make_hard_clip(w: int, h: int, nframes: int, seed: int, start: int = 0, noise: int = 3) -> np.ndarray:
    """Content that defeats the encoder's cheap paths: a textured picture panning over the WHOLE frame by
    fractional samples ((1.25, 0.75) luma samples per frame, bilinear resampling) with fresh noise of
    +-`noise` on every sample of every frame.  No macroblock is static, no vector is integer, no residual is
    empty: the sub-sample search, the transform and the entropy coder all run on every macroblock."""
    rng = np.random.default_rng(seed)
    pad = 64 + int(1.25 * (start + nframes)) + 2
    # band-limited texture: white noise smoothed by a separable box filter, plus a coarse pattern for large structures
    tex = rng.integers(0, 256, size=(h + pad, w + pad)).astype(np.float32)
    for ax in (0, 1):
        tex = (np.roll(tex, 1, ax) + tex + np.roll(tex, -1, ax) + np.roll(tex, 2, ax)) * 0.25
    yy, xx = np.mgrid[0:h + pad, 0:w + pad]
    tex = 0.6 * tex + 0.4 * (128 + 90 * np.sin(xx / 23.0) * np.cos(yy / 17.0))
    tex = np.clip(tex, 16, 235).astype(np.float32)
    cu = (128 + 40 * np.sin(xx[::2, ::2] / 61.0)).astype(np.float32)
    cv = (128 + 40 * np.cos(yy[::2, ::2] / 47.0)).astype(np.float32)
    cw, ch = (w + 1) // 2, (h + 1) // 2
    out = np.empty((nframes, frame_bytes(w, h)), np.uint8)

    def sample(plane, fx, fy, pw, ph):
        ix, iy = int(np.floor(fx)), int(np.floor(fy))
        ax, ay = np.float32(fx - ix), np.float32(fy - iy)
        p00 = plane[iy:iy + ph, ix:ix + pw]; p01 = plane[iy:iy + ph, ix + 1:ix + 1 + pw]
        p10 = plane[iy + 1:iy + 1 + ph, ix:ix + pw]; p11 = plane[iy + 1:iy + 1 + ph, ix + 1:ix + 1 + pw]
        return (1 - ax) * (1 - ay) * p00 + ax * (1 - ay) * p01 + (1 - ax) * ay * p10 + ax * ay * p11

    for i in range(nframes):
        n = start + i
        fx, fy = 1.25 * n, 0.75 * n
        Y = sample(tex, fx, fy, w, h) + rng.integers(-noise, noise + 1, size=(h, w)).astype(np.float32)
        U = sample(cu, fx / 2, fy / 2, cw, ch)
        V = sample(cv, fx / 2, fy / 2, cw, ch)
        f = out[i]
        f[: w * h] = np.clip(Y + 0.5, 0, 255).astype(np.uint8).ravel()
        f[w * h: w * h + cw * ch] = np.clip(U + 0.5, 0, 255).astype(np.uint8).ravel()
        f[w * h + cw * ch:] = np.clip(V + 0.5, 0, 255).astype(np.uint8).ravel()
    return out
