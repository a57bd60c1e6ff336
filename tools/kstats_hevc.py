"""Developer script: per-kernel CUDA-event times (single stream) of the HEVC path on the GPU."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_codec_pipeline_b200 import api, synth


def run(w, h, gops, gop, qp=27):
    a = synth.make_clip(w, h, gop, seed=1080)
    clip = np.concatenate([a] * gops, axis=0)
    n = clip.shape[0]
    p = api.default_params(w, h, codec=1, gop=gop, qp_i=qp - 2, qp_p=qp, slices=0)
    with api.Session(p, n) as s:
        s.upload(clip); s.encode()
        s.profile(True)
        s.upload(clip); ms = s.encode()
        st = s.kernel_stats()
        size = len(s.download()["stream"])
    print("hevc %dx%d gops=%d total=%.2f ms single-stream, %d frames, %.1f kbit/frame" % (w, h, gops, ms, n, size * 8 / n / 1000))
    for k, v in st.items():
        if v["launches"]:
            print("   %-14s %8.2f ms total  %8.1f us x %d" % (k, v["ms"], 1000 * v["ms"] / v["launches"], v["launches"]))


if __name__ == "__main__":
    g = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    run(1920, 1080, g, 60)
