"""Developer script: CUDA path vs oracle vs FFmpeg decoder on one small clip (run under gpurun)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_codec_pipeline_b200 import api, synth, arbiter
from oracle import pyoracle

def run(w, h, n, gop, slices, idc, qp, seed=7):
    clip = synth.make_clip(w, h, n, seed=seed)
    op = pyoracle.make_params(w, h, gop=gop, qp_i=qp - 2, qp_p=qp, slices=slices, deblock_idc=idc)
    ref = pyoracle.encode(op, clip, want_dump=True)
    p = api.default_params(w, h, gop=gop, qp_i=qp - 2, qp_p=qp, slices=slices, deblock_idc=idc, debug=1)
    with api.Session(p, n) as s:
        s.upload(clip)
        ms = s.encode()
        got = s.download(want_recon=True)
        dbg = s.debug_mbs()
    tag = "%dx%d n=%d gop=%d sl=%d idc=%d qp=%d" % (w, h, n, gop, slices, idc, qp)
    ok = True
    for k in ("mv_prepass", "mv_final", "mb_type", "cbp"):
        a, b = dbg[k], ref["dump"][k]
        if k == "mv_prepass":
            # oracle leaves IDR frames at zero; the device buffer too (memset at create)
            pass
        bad = np.argwhere(a != b)
        if len(bad):
            ok = False
            print("  MISMATCH", k, len(bad), "first", bad[0], "gpu", a[tuple(bad[0])], "oracle", b[tuple(bad[0])])
    rec_bad = [i for i in range(n) if not np.array_equal(got["recon"][i], ref["recon"][i])]
    if rec_bad:
        ok = False
        i = rec_bad[0]
        d = np.argwhere(got["recon"][i] != ref["recon"][i])
        print("  RECON mismatch frames", rec_bad[:8], "first byte", d[0], "count", len(d))
    gs, os_ = got["stream"].tobytes(), ref["stream"]
    if gs != os_:
        ok = False
        m = next((i for i in range(min(len(gs), len(os_))) if gs[i] != os_[i]), min(len(gs), len(os_)))
        print("  STREAM mismatch: len gpu %d oracle %d first diff at %d" % (len(gs), len(os_), m),
              "frame sizes gpu", [x[1] for x in got["info"]][:6], "oracle", [x[1] for x in ref["info"]][:6])
    # arbiter: decoder must reproduce the GPU's own recon
    dec = arbiter.decode_annexb(gs)
    dec_ok = len(dec) == n and all(
        np.array_equal(np.concatenate([pl.ravel() for pl in dec[i]]), got["recon"][i]) for i in range(n))
    print("%s  %s: oracle-identical=%s decoder-bitexact=%s bytes=%d gpu_ms=%.2f" %
          ("PASS" if ok and dec_ok else "FAIL", tag, ok, dec_ok, len(gs), ms))
    return ok and dec_ok

if __name__ == "__main__":
    print(api.version(), "devices", api.device_count())
    cases = [(64, 48, 3, 60, 1, 1, 26), (64, 48, 3, 60, 1, 0, 26), (320, 180, 8, 4, 3, 0, 30),
             (320, 180, 8, 4, 3, 2, 18), (640, 360, 6, 60, 1, 0, 40), (176, 144, 10, 5, 2, 0, 12),
             (176, 144, 10, 5, 9, 0, 51), (1920, 1080, 4, 60, 1, 0, 27)]
    if len(sys.argv) > 1:
        cases = cases[: int(sys.argv[1])]
    res = [run(*c) for c in cases]
    print("ALL PASS" if all(res) else "SOME FAILED")
    sys.exit(0 if all(res) else 1)
