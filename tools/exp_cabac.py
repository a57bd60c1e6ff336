"""Developer script: step time of the 1080p CABAC configuration under experiment knobs (GPU)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_codec_pipeline_b200 import api, synth
w, h, gop, gops = 1920, 1080, 60, 32
a = synth.make_clip(w, h, gop, seed=1080)
clip = np.concatenate([a] * gops, axis=0)
n = clip.shape[0]
def run(tag, **kw):
    p = api.default_params(w, h, gop=gop, qp_i=25, qp_p=27, **kw)
    with api.Session(p, n) as s:
        s.upload(clip)
        for _ in range(2): s.encode()
        ms = [s.encode() for _ in range(3)]
    print(tag, "%.1f ms/step -> %.0f fps" % (np.mean(ms), n / np.mean(ms) * 1000), flush=True)
run(sys.argv[1] if len(sys.argv) > 1 else "cabac", entropy=1, slices=int(sys.argv[2]) if len(sys.argv) > 2 else 0)
