"""Quality / rate report of the B200 encoder on the synthetic S1080 clip:
QP sweep {22,27,32,37} for both entropy coders and one bitrate-target run at 10 Mb/s.  For every
point: bitrate, PSNR-Y/U/V and SSIM-Y of the DECODED stream (FFmpeg h264 / hevc decoder) against the source,
decoder-vs-encoder-reconstruction bit-exactness, and the MP4 verify result.
The libx264 side of the north-star's "within 0.5 dB at matched bitrate" cannot be produced in this
image (no ffmpeg / libx264): the JSON says so instead of inventing numbers.

    python tools/quality.py [--frames 120] [--out profiles/r01_quality.json]     (needs the GPU)
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_codec_pipeline_b200 import api, arbiter, synth  # noqa: E402


def ssim_plane(a, b):
    """Mean SSIM over 8x8 non-overlapping windows (C1/C2 of the original paper, 8-bit)."""
    a = a.astype(np.float64); b = b.astype(np.float64)
    h, w = a.shape[0] // 8 * 8, a.shape[1] // 8 * 8
    A = a[:h, :w].reshape(h // 8, 8, w // 8, 8).transpose(0, 2, 1, 3).reshape(-1, 64)
    B = b[:h, :w].reshape(h // 8, 8, w // 8, 8).transpose(0, 2, 1, 3).reshape(-1, 64)
    ma, mb = A.mean(1), B.mean(1)
    va, vb = A.var(1), B.var(1)
    cov = ((A - ma[:, None]) * (B - mb[:, None])).mean(1)
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    return float((((2 * ma * mb + c1) * (2 * cov + c2)) / ((ma ** 2 + mb ** 2 + c1) * (va + vb + c2))).mean())


def measure(p, clip, w, h, fps, label):
    n = clip.shape[0]
    got = api.encode_frames(p, clip, want_recon=True)
    stream = got["stream"].tobytes()
    dec = arbiter.decode_annexb_hevc(stream, threads=8) if p.codec == 1 else arbiter.decode_annexb(stream, threads=8)
    exact = len(dec) == n and all(np.array_equal(np.concatenate([pl.ravel() for pl in dec[i]]), got["recon"][i]) for i in range(n))
    ps = [[], [], []]
    ss = []
    for i in range(0, n, max(1, n // 30)):
        src = synth.split_planes(clip[i], w, h)
        for k in range(3):
            ps[k].append(arbiter.psnr(dec[i][k], src[k]))
        ss.append(ssim_plane(dec[i][0], src[0]))
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "o.mp4")
        api.mux_mp4(p, got["stream"], got["info"], path)
        try:
            api.verify(path)
            ver = True
        except api.VcpencError:
            ver = False
    return {"point": label, "bytes": len(stream), "kbps": round(len(stream) * 8 * fps / n / 1000.0, 1),
            "psnr_y": round(float(np.mean(ps[0])), 3), "psnr_u": round(float(np.mean(ps[1])), 3),
            "psnr_v": round(float(np.mean(ps[2])), 3), "ssim_y": round(float(np.mean(ss)), 5),
            "decoder_bit_exact": bool(exact), "verify": ver, "qp_first_p": got["info"][1][3] if n > 1 else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=120)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_quality.json"))
    a = ap.parse_args()
    w, h, fps, gop = 1920, 1080, 30, 60
    clip = synth.make_clip(w, h, a.frames, seed=1080)
    rows = []
    for ent in (0, 1):
        for qp in (22, 27, 32, 37):
            p = api.default_params(w, h, fps=fps, gop=gop, qp_i=qp - 2, qp_p=qp, entropy=ent, slices=0)
            rows.append(measure(p, clip, w, h, fps, "%s qp%d" % ("cabac" if ent else "cavlc", qp)))
            print(rows[-1], flush=True)
    for t8 in (1,):       # High profile: CABAC + 8x8 transform, the default of the libx264 / nvenc presets
        for qp in (22, 27, 32, 37):
            p = api.default_params(w, h, fps=fps, gop=gop, qp_i=qp - 2, qp_p=qp, entropy=1, transform8x8=t8, slices=0)
            rows.append(measure(p, clip, w, h, fps, "high qp%d" % qp))
            print(rows[-1], flush=True)
    for idc in (0, 1):    # HEVC (h265-* presets), with and without in-loop deblocking
        for qp in (22, 27, 32, 37):
            p = api.default_params(w, h, fps=fps, gop=gop, qp_i=qp - 2, qp_p=qp, codec=1, deblock_idc=idc, slices=0)
            rows.append(measure(p, clip, w, h, fps, "hevc%s qp%d" % (" no-deblock" if idc else "", qp)))
            print(rows[-1], flush=True)
    p = api.default_params(w, h, fps=fps, gop=gop, rc_mode=1, bitrate=8_000_000, codec=1, slices=0)
    rows.append(measure(p, clip, w, h, fps, "hevc 8Mb/s target (h265-nvenc preset)"))
    print(rows[-1], flush=True)
    for ent in (0, 1):
        p = api.default_params(w, h, fps=fps, gop=gop, rc_mode=1, bitrate=10_000_000, entropy=ent, slices=0)
        rows.append(measure(p, clip, w, h, fps, "%s 10Mb/s target" % ("cabac" if ent else "cavlc")))
        print(rows[-1], flush=True)
    rep = {"clip": "S1080 stand-in: synth.make_clip(1920,1080,%d,seed=1080), 30 fps, GOP 60" % a.frames,
           "encoder": api.version(), "points": rows,
           "libx264": None, "libx264_reason": "ffmpeg/libx264 not present in the image (SURVEY 0.3): matched-bitrate delta unmeasured"}
    json.dump(rep, open(a.out, "w"), indent=1)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
