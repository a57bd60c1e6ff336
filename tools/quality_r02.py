"""Round-2 quality report: what the tools added in round 2 buy, as Bjontegaard rate differences on the HARD clip
(full-frame fractional pan + noise, synth.make_hard_clip) and on the standard clip, 1080p, GOP 60.

Curves (QP 22/27/32/37 each; every point decoded by the FFmpeg decoder and compared with the encoder's reconstruction):
  h264 high effort1   High profile CABAC, the `h264-cpu` tools as parsed (round-1 encoder = this curve)
  h264 high effort0   `-preset` fast tiers: no quarter-sample step
  hevc subpel1        round-1 HEVC: half-sample motion
  hevc subpel2        round 2: quarter-sample motion, candidates ranked by the half-sample proxy (preset default)
  hevc subpel3        round 2: quarter-sample motion, candidates ranked by the exact prediction
plus one VBV point (the hard clip at QP 22/24 capped with -maxrate / -bufsize; the decoder-buffer trace of the stream).
libx264 / libx265 stay null: neither exists in the image.

    python tools/quality_r02.py [--frames 120] [--out profiles/r02_quality.json]      (needs the GPU)
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from video_codec_pipeline_b200 import api, synth  # noqa: E402
from quality import measure  # noqa: E402


def bd_rate(ref, test):
    """Bjontegaard delta rate (%) of `test` against `ref`: [(kbps, psnr)] x 4, cubic fit of log-rate over PSNR."""
    r1, p1 = np.log([x[0] for x in ref]), np.array([x[1] for x in ref])
    r2, p2 = np.log([x[0] for x in test]), np.array([x[1] for x in test])
    c1, c2 = np.polyfit(p1, r1, 3), np.polyfit(p2, r2, 3)
    lo, hi = max(p1.min(), p2.min()), min(p1.max(), p2.max())
    if hi <= lo:
        return None
    i1, i2 = np.polyint(c1), np.polyint(c2)
    avg = ((np.polyval(i2, hi) - np.polyval(i2, lo)) - (np.polyval(i1, hi) - np.polyval(i1, lo))) / (hi - lo)
    return round(float((np.exp(avg) - 1) * 100), 2)


def vbv_min_fullness(bits, rate, buf, start):
    f, lo = start, float("inf")
    for b in bits:
        f = min(buf, f + rate) - b
        lo = min(lo, f)
    return lo


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=120)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_quality.json"))
    a = ap.parse_args()
    w, h, fps, gop = 1920, 1080, 30, 60
    rep = {"encoder": api.version(), "size": "1920x1080, %d frames, 30 fps, GOP 60" % a.frames, "clips": {},
           "libx264": None, "libx264_reason": "ffmpeg/libx264/libx265 not present in the image (SURVEY 0.3): matched-bitrate delta unmeasured"}
    for cname, clip in (("hard", synth.make_hard_clip(w, h, a.frames, seed=7)), ("standard", synth.make_clip(w, h, a.frames, seed=1080))):
        curves = {}
        cfgs = (("h264 high effort1", dict(entropy=1, transform8x8=1, effort=1)), ("h264 high effort0", dict(entropy=1, transform8x8=1, effort=0)),
                ("hevc subpel1", dict(codec=1, hevc_subpel=1)), ("hevc subpel2", dict(codec=1, hevc_subpel=2)), ("hevc subpel3", dict(codec=1, hevc_subpel=3)))
        for label, kw in cfgs:
            rows = []
            for qp in (22, 27, 32, 37):
                p = api.default_params(w, h, fps=fps, gop=gop, qp_i=qp - 2, qp_p=qp, slices=0, **kw)
                rows.append(measure(p, clip, w, h, fps, "%s qp%d" % (label, qp)))
                print(cname, rows[-1], flush=True)
            curves[label] = rows
        pts = {k: [(r["kbps"], r["psnr_y"]) for r in v] for k, v in curves.items()}
        rep["clips"][cname] = {
            "curves": curves,
            "bd_rate_percent_psnr_y": {
                "h264 effort0 (fast tiers) vs effort1": bd_rate(pts["h264 high effort1"], pts["h264 high effort0"]),
                "hevc subpel2 (round 2 default) vs subpel1 (round 1)": bd_rate(pts["hevc subpel1"], pts["hevc subpel2"]),
                "hevc subpel3 (exact ranking) vs subpel2 (proxy ranking)": bd_rate(pts["hevc subpel2"], pts["hevc subpel3"]),
                "hevc subpel2 vs h264 high": bd_rate(pts["h264 high effort1"], pts["hevc subpel2"]),
                "hevc subpel1 (round 1) vs h264 high": bd_rate(pts["h264 high effort1"], pts["hevc subpel1"]),
            },
            "all_points_decoder_bit_exact": all(r["decoder_bit_exact"] for v in curves.values() for r in v),
        }
        print(cname, rep["clips"][cname]["bd_rate_percent_psnr_y"], flush=True)
        if cname == "hard":
            # VBV: the clip at CRF-like QP 21/24 needs more than the cap allows
            free = api.encode_frames(api.default_params(w, h, fps=fps, gop=gop, qp_i=21, qp_p=24, entropy=1, transform8x8=1, slices=0), clip)
            bits0 = [x[1] * 8 for x in free["info"]]
            natural = sum(bits0) * fps / len(bits0)
            maxrate, bufsize = int(0.6 * natural), int(0.6 * natural)
            capped = api.encode_frames(api.default_params(w, h, fps=fps, gop=gop, qp_i=21, qp_p=24, entropy=1, transform8x8=1, slices=0,
                                                          maxrate=maxrate, bufsize=bufsize), clip)
            bits = [x[1] * 8 for x in capped["info"]]
            rep["vbv"] = {"clip": "hard", "qp_i_p": [21, 24], "uncapped_kbps": round(natural / 1000, 1), "maxrate": maxrate, "bufsize": bufsize,
                          "capped_kbps": round(sum(bits) * fps / len(bits) / 1000, 1),
                          "min_fullness_of_bufsize_uncapped": round(vbv_min_fullness(bits0, maxrate / fps, bufsize, 0.9 * bufsize) / bufsize, 3),
                          "min_fullness_of_bufsize_capped": round(vbv_min_fullness(bits, maxrate / fps, bufsize, 0.9 * bufsize) / bufsize, 3),
                          "qp_range_capped": [min(x[3] for x in capped["info"]), max(x[3] for x in capped["info"])]}
            print(rep["vbv"], flush=True)
    json.dump(rep, open(a.out, "w"), indent=1)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
