"""Generates tests/golden/h264_golden.json and tests/golden/hevc_golden.json.

There are no golden vectors in the reference (it has no tests), so these
are produced by the CPU oracle on seeded synthetic clips and *validated at generation time* by
the FFmpeg h264 decoder: a case is only written if the decoder reproduces the oracle's
reconstruction bit-exactly.  Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import CASES  # noqa: E402
from oracle import pyoracle  # noqa: E402
from video_codec_pipeline_b200 import arbiter, synth  # noqa: E402

out = []
for (w, h, n, gop, sl, idc, qp, ent, t8) in [c + (0, 0) for c in CASES] + [c + (1, 0) for c in CASES] + \
        [c + (0, 1) for c in CASES[2:]] + [c + (1, 1) for c in CASES[2:]]:
    clip = synth.make_clip(w, h, n, seed=1000 + w + qp)
    p = pyoracle.make_params(w, h, gop=gop, qp_i=max(0, qp - 2), qp_p=qp, slices=sl, deblock_idc=idc, entropy=ent, transform8x8=t8)
    r = pyoracle.encode(p, clip)
    dec = arbiter.decode_annexb(r["stream"])
    assert len(dec) == n
    for i in range(n):
        flat = np.concatenate([pl.ravel() for pl in dec[i]])
        assert np.array_equal(flat, r["recon"][i]), (w, h, i)
    out.append({
        "w": w, "h": h, "frames": n, "gop": gop, "slices": sl, "deblock_idc": idc, "qp": qp, "entropy": ent, "transform8x8": t8,
        "seed": 1000 + w + qp,
        "clip_sha256": hashlib.sha256(clip.tobytes()).hexdigest(),
        "stream_bytes": len(r["stream"]),
        "stream_sha256": hashlib.sha256(r["stream"]).hexdigest(),
        "recon_sha256": hashlib.sha256(r["recon"].tobytes()).hexdigest(),
        "frame_sizes": [x[1] for x in r["info"]],
    })
    print("ok", w, h, n, gop, sl, idc, qp, ent, t8, len(r["stream"]))
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "h264_golden.json"), "w") as f:
    json.dump(out, f, indent=1)

# ---- HEVC (oracle/hevc_oracle.inc.c), validated by the FFmpeg hevc decoder; both motion precisions -------------
out = []
for (w, h, n, gop, sl, idc, qp) in CASES:
    for sub in (0, 1):
        clip = synth.make_clip(w, h, n, seed=1000 + w + qp)
        p = pyoracle.make_params(w, h, codec=1, gop=gop, qp_i=max(0, qp - 2), qp_p=qp, slices=sl, deblock_idc=idc, hevc_subpel=sub)
        r = pyoracle.encode_hevc(p, clip)
        dec = arbiter.decode_annexb_hevc(r["stream"])
        assert len(dec) == n
        for i in range(n):
            assert np.array_equal(np.concatenate([pl.ravel() for pl in dec[i]]), r["recon"][i]), (w, h, i)
        out.append({
            "w": w, "h": h, "frames": n, "gop": gop, "slices": sl, "deblock_idc": idc, "qp": qp, "hevc_subpel": sub,
            "seed": 1000 + w + qp,
            "clip_sha256": hashlib.sha256(clip.tobytes()).hexdigest(),
            "stream_bytes": len(r["stream"]),
            "stream_sha256": hashlib.sha256(r["stream"]).hexdigest(),
            "recon_sha256": hashlib.sha256(r["recon"].tobytes()).hexdigest(),
            "frame_sizes": [x[1] for x in r["info"]],
        })
        print("ok hevc", w, h, n, gop, sl, idc, qp, sub, len(r["stream"]))
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "hevc_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
