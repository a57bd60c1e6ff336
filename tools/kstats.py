"""Developer script: per-kernel CUDA-event times for a few encoder configurations (GPU)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_codec_pipeline_b200 import api, synth

def run(w, h, gops, gop, slices, idc, qp=27, tag=""):
    a = synth.make_clip(w, h, gop, seed=1080)
    clip = np.concatenate([a] * gops, axis=0)
    n = clip.shape[0]
    p = api.default_params(w, h, gop=gop, qp_i=qp - 2, qp_p=qp, slices=slices, deblock_idc=idc)
    with api.Session(p, n) as s:
        s.upload(clip); s.encode()
        s.profile(True)
        s.upload(clip); ms = s.encode()
        st = s.kernel_stats()
    print("%s %dx%d gops=%d slices=%d idc=%d total=%.2fms" % (tag, w, h, gops, slices, idc, ms))
    print("   ", {k: "%.1fus x%d" % (1000 * v["ms"] / v["launches"], v["launches"]) for k, v in st.items() if v["launches"]})

if __name__ == "__main__":
    g = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    if len(sys.argv) > 2 and sys.argv[2] == "indep":
        run(1920, 1080, g, 8, 68, 2, tag="rows-independent")
        sys.exit(0)
    run(1920, 1080, g, 60, 1, 0, tag="base")
    run(1920, 1080, g, 60, 68, 2, tag="rows-independent")
    run(1920, 1080, g, 60, 4, 2, tag="4 slices idc2")
    run(1920, 1080, g, 60, 1, 0, qp=40, tag="qp40")
