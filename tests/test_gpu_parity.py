"""Parity tests proper (run on the B200): everything goes through the C-ABI (libvcpenc.so).

Bar: bit-exact.  (1) the CUDA bitstream, reconstruction and per-macroblock decisions equal the
CPU oracle's on the same seeded inputs; (2) the committed golden hashes; (3) at BASELINE.json's
full sizes, size-independent properties: the FFmpeg decoder reproduces the encoder's own
reconstruction exactly, encodes are deterministic, and a GOP-sharded encode concatenates to the
unsharded stream."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import CASES
from video_codec_pipeline_b200 import api, arbiter, synth
from video_codec_pipeline_b200.shard import gop_ranges

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "h264_golden.json")))
GOLD_HEVC = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hevc_golden.json")))


def _flat(planes):
    return np.concatenate([pl.ravel() for pl in planes])


@pytest.mark.parametrize("entropy,t8", [(0, 0), (1, 0), (0, 1), (1, 1)], ids=["cavlc", "cabac", "cavlc-high", "cabac-high"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d_n%d_g%d_s%d_d%d_q%d" % c)
def test_cuda_equals_oracle(built, case, entropy, t8):
    from oracle import pyoracle
    w, h, n, gop, sl, idc, qp = case
    clip = synth.make_clip(w, h, n, seed=1000 + w + qp)
    ref = pyoracle.encode(pyoracle.make_params(w, h, gop=gop, qp_i=max(0, qp - 2), qp_p=qp, slices=sl, deblock_idc=idc,
                                               entropy=entropy, transform8x8=t8), clip, want_dump=True)
    p = api.default_params(w, h, gop=gop, qp_i=max(0, qp - 2), qp_p=qp, slices=sl, deblock_idc=idc, debug=1, entropy=entropy,
                           transform8x8=t8)
    with api.Session(p, n) as s:
        s.upload(clip)
        s.encode()
        got = s.download(want_recon=True)
        dbg = s.debug_mbs()
    for k in ("mv_prepass", "mv_final", "mb_type", "cbp"):        # K2a, K2b, mbinfo, K3
        assert np.array_equal(dbg[k], ref["dump"][k]), k
    assert np.array_equal(got["recon"], ref["recon"])             # K3 + K4
    assert got["stream"].tobytes() == ref["stream"]               # K5 + NAL packing
    assert [x[1] for x in got["info"]] == [x[1] for x in ref["info"]]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d_n%d_g%d_s%d_d%d_q%d" % c)
def test_hevc_cuda_equals_oracle(built, case):
    """HEVC path (k6_hevc.cu + the shared motion search and arithmetic coder) against oracle/hevc_oracle.inc.c:
    pre-pass vectors, reconstruction and Annex-B bytes identical; the oracle itself is pinned by the FFmpeg hevc
    decoder (tests/test_oracle.py)."""
    from oracle import pyoracle
    w, h, n, gop, sl, idc, qp = case                       # idc 1: in-loop deblocking off
    clip = synth.make_clip(w, h, n, seed=1000 + w + qp)
    ref = pyoracle.encode_hevc(pyoracle.make_params(w, h, codec=1, gop=gop, qp_i=max(0, qp - 2), qp_p=qp, slices=sl, deblock_idc=idc), clip)
    p = api.default_params(w, h, codec=1, gop=gop, qp_i=max(0, qp - 2), qp_p=qp, slices=sl, debug=1, deblock_idc=idc)
    with api.Session(p, n) as s:
        s.upload(clip)
        s.encode()
        got = s.download(want_recon=True)
    bad = [i for i in range(n) if not np.array_equal(got["recon"][i], ref["recon"][i])]
    assert not bad, "recon differs in frames %s" % bad
    assert [x[1] for x in got["info"]] == [x[1] for x in ref["info"]]
    assert got["stream"].tobytes() == ref["stream"]


def test_hevc_full_size_and_bitrate_mode(built):
    """1080p HEVC: the FFmpeg hevc decoder reproduces the encoder's reconstruction; -b:v mode equals the oracle."""
    from oracle import pyoracle
    w, h, n, gop = 1920, 1080, 8, 4
    clip = synth.make_clip(w, h, n, seed=77)
    got = api.encode_frames(api.default_params(w, h, codec=1, gop=gop, qp_i=26, qp_p=28, slices=0, debug=1), clip, want_recon=True)
    if arbiter.available():
        dec = arbiter.decode_annexb_hevc(got["stream"].tobytes())
        assert len(dec) == n
        for i in range(n):
            assert np.array_equal(_flat(dec[i]), got["recon"][i]), "frame %d" % i
        assert arbiter.psnr(dec[n - 1][0], synth.split_planes(clip[n - 1], w, h)[0]) > 30
    w, h, n, gop = 320, 240, 24, 12
    clip = synth.make_clip(w, h, n, seed=78)
    kw = dict(codec=1, gop=gop, slices=2, rc_mode=1, bitrate=600_000, fps_num=30, fps_den=1)
    ref = pyoracle.encode_hevc(pyoracle.make_params(w, h, **kw), clip)
    got = api.encode_frames(api.default_params(w, h, **kw), clip)
    assert [x[3] for x in got["info"]] == [x[3] for x in ref["info"]]      # the QP path of the rate control
    assert got["stream"].tobytes() == ref["stream"]
    assert len({x[3] for x in ref["info"]}) > 1


@pytest.mark.parametrize("g", GOLD, ids=lambda g: "%dx%d_q%d_s%d_e%d_t%d" % (g["w"], g["h"], g["qp"], g["slices"], g.get("entropy", 0), g.get("transform8x8", 0)))
def test_cuda_matches_golden(built, g):
    clip = synth.make_clip(g["w"], g["h"], g["frames"], seed=g["seed"])
    p = api.default_params(g["w"], g["h"], gop=g["gop"], qp_i=max(0, g["qp"] - 2), qp_p=g["qp"],
                           slices=g["slices"], deblock_idc=g["deblock_idc"], entropy=g.get("entropy", 0),
                           transform8x8=g.get("transform8x8", 0))
    got = api.encode_frames(p, clip, want_recon=True)             # host buffers in, host buffers out
    assert [x[1] for x in got["info"]] == g["frame_sizes"]
    assert hashlib.sha256(got["stream"].tobytes()).hexdigest() == g["stream_sha256"]
    assert hashlib.sha256(got["recon"].tobytes()).hexdigest() == g["recon_sha256"]


@pytest.mark.parametrize("g", GOLD_HEVC, ids=lambda g: "hevc_%dx%d_q%d_s%d_d%d_h%d" % (g["w"], g["h"], g["qp"], g["slices"], g["deblock_idc"], g["hevc_subpel"]))
def test_hevc_cuda_matches_golden(built, g):
    clip = synth.make_clip(g["w"], g["h"], g["frames"], seed=g["seed"])
    p = api.default_params(g["w"], g["h"], codec=1, gop=g["gop"], qp_i=max(0, g["qp"] - 2), qp_p=g["qp"], slices=g["slices"],
                           deblock_idc=g["deblock_idc"], hevc_subpel=g["hevc_subpel"], debug=1)
    got = api.encode_frames(p, clip, want_recon=True)
    assert [x[1] for x in got["info"]] == g["frame_sizes"]
    assert hashlib.sha256(got["stream"].tobytes()).hexdigest() == g["stream_sha256"]
    assert hashlib.sha256(got["recon"].tobytes()).hexdigest() == g["recon_sha256"]


@pytest.mark.parametrize("w,h,n,gop,sl,kw", [(1920, 1080, 24, 8, 1, {}), (1920, 1080, 12, 6, 4, {}), (3840, 2160, 6, 3, 1, {}),
                                             (1920, 1080, 12, 6, 0, dict(entropy=1, transform8x8=1)),
                                             (3840, 2160, 6, 3, 0, dict(entropy=1, transform8x8=1))],
                         ids=["1080p", "1080p-4slices", "4k", "1080p-high-cabac", "4k-high-cabac"])
def test_full_size_decoder_reproduces_recon(built, w, h, n, gop, sl, kw):
    if not arbiter.available():
        pytest.skip("bundled FFmpeg decoder not present")
    clip = synth.make_clip(w, h, n, seed=w)
    p = api.default_params(w, h, gop=gop, qp_i=25, qp_p=27, slices=sl, debug=1, **kw)
    got = api.encode_frames(p, clip, want_recon=True)
    dec = arbiter.decode_annexb(got["stream"].tobytes(), threads=8)
    assert len(dec) == n
    for i in range(n):
        assert np.array_equal(_flat(dec[i]), got["recon"][i]), "frame %d" % i
    y = synth.split_planes(clip[n - 1], w, h)[0]
    assert arbiter.psnr(dec[n - 1][0], y) > 30
    # determinism: same input, same bytes
    again = api.encode_frames(p, clip)
    assert again["stream"].tobytes() == got["stream"].tobytes()
    # GOP sharding: encode the ranges separately (as 2 and 3 GPUs would) and concatenate
    for world in (2, 3):
        parts = []
        for f0, cnt, g0 in gop_ranges(n, gop, world):
            if cnt:
                q = api.default_params(w, h, gop=gop, qp_i=25, qp_p=27, slices=sl, first_gop=g0, **kw)
                parts.append(api.encode_frames(q, clip[f0:f0 + cnt])["stream"].tobytes())
        assert b"".join(parts) == got["stream"].tobytes()


def test_bitrate_mode_equals_oracle(built):
    """Rate control runs on the device (rc_update_kernel) with the feedback delay of the side-stream
    entropy coder; the oracle mirrors it, so bytes and per-picture QPs must agree."""
    from oracle import pyoracle
    w, h, n, fps = 320, 192, 50, 25
    clip = synth.make_clip(w, h, n, seed=12)
    for br, gop, sl in ((500_000, 20, 1), (2_000_000, 24, 3)):
        ref = pyoracle.encode(pyoracle.make_params(w, h, fps=fps, gop=gop, slices=sl, rc_mode=1, bitrate=br), clip)
        got = api.encode_frames(api.default_params(w, h, fps=fps, gop=gop, slices=sl, rc_mode=1, bitrate=br), clip,
                                want_recon=True)
        assert [x[3] for x in got["info"]] == [x[3] for x in ref["info"]]
        assert got["stream"].tobytes() == ref["stream"]
        assert np.array_equal(got["recon"], ref["recon"])


def test_vbv_equals_oracle(built):
    """-maxrate / -bufsize: the per-GOP buffer model runs in rc_update_kernel (vcp_algo.h: vcp_rc_picture, shared with the
    oracle); per-picture QPs, bytes and reconstruction must agree in constant-QP mode (CAVLC, High CABAC, HEVC) and
    together with -b:v, and the cap must actually move the QPs."""
    w, h, n, fps, gop = 320, 192, 60, 24, 24
    clip = synth.make_hard_clip(w, h, n, seed=5)
    from oracle import pyoracle
    for kw in (dict(entropy=0), dict(entropy=1, transform8x8=1, slices=2), dict(entropy=0, rc_mode=1, bitrate=900_000)):
        got = _same_as_oracle(w, h, clip, fps=fps, gop=gop, qp_i=21, qp_p=24, maxrate=300_000, bufsize=300_000, **kw)
        qps = [x[3] for x in got["info"]]
        assert len(set(qps[1:gop])) > 2, (kw, qps)
    kw = dict(codec=1, entropy=1, slices=2, fps=fps, gop=gop, qp_i=24, qp_p=27, maxrate=300_000, bufsize=300_000)
    ref = pyoracle.encode_hevc(pyoracle.make_params(w, h, **kw), clip)
    got = api.encode_frames(api.default_params(w, h, **kw), clip, want_recon=True)
    assert [x[3] for x in got["info"]] == [x[3] for x in ref["info"]]
    assert got["stream"].tobytes() == ref["stream"] and np.array_equal(got["recon"], ref["recon"])
    assert len({x[3] for x in ref["info"][1:gop]}) > 2


@pytest.mark.parametrize("fmt", [1, 2, 3, 4, 5])
def test_k1_input_formats_equal_oracle(built, fmt):
    """K1 front stages (nv12 / rgb24 / yuv444p / yuv422p / bgr24, with and without scaling)."""
    from oracle import pyoracle
    rng = np.random.default_rng(100 + fmt)
    for (iw, ih, ow, oh) in ((96, 64, 96, 64), (130, 98, 64, 48), (64, 48, 208, 114)):
        kw = dict(gop=2, in_fmt=fmt, in_width=iw, in_height=ih)
        nb = api.in_frame_bytes(api.default_params(ow, oh, **kw))
        # smooth content plus noise so that both prediction paths are exercised
        base = np.linspace(30, 220, nb)[None, :] + rng.integers(-20, 21, (3, nb))
        frames = np.clip(base, 0, 255).astype(np.uint8)
        ref = pyoracle.encode(pyoracle.make_params(ow, oh, **kw), frames)
        got = api.encode_frames(api.default_params(ow, oh, **kw), frames, want_recon=True)
        assert got["stream"].tobytes() == ref["stream"], (fmt, iw, ih, ow, oh)
        assert np.array_equal(got["recon"], ref["recon"])


def _same_as_oracle(w, h, clip, **kw):
    from oracle import pyoracle
    ref = pyoracle.encode(pyoracle.make_params(w, h, **kw), clip)
    got = api.encode_frames(api.default_params(w, h, **kw), clip, want_recon=True)
    assert [x[3] for x in got["info"]] == [x[3] for x in ref["info"]], kw
    assert got["stream"].tobytes() == ref["stream"], kw
    assert np.array_equal(got["recon"], ref["recon"]), kw
    return got


def test_breadth_against_oracle(built):
    """Corners the headline configurations do not reach: long GOPs across several CABAC batches, cropped
    sizes (1366x768), fast pans that push vectors to the border clamp, flat pictures, rate control with
    High profile and slices, nv12 at 1080p, and a session reused for different content."""
    rng = np.random.default_rng(4)
    # GOP 250-style long GOP (62 pictures in one GOP = 8 CABAC batches) + a ragged second GOP
    w, h = 176, 144
    _same_as_oracle(w, h, synth.make_clip(w, h, 70, seed=2), gop=62, qp_i=28, qp_p=30, entropy=1, transform8x8=1, slices=2)
    # width / height not multiples of 16: frame cropping in the SPS, padded macroblocks
    w, h = 1366, 768
    g = _same_as_oracle(w, h, synth.make_clip(w, h, 4, seed=3), gop=4, qp_i=26, qp_p=28, entropy=1, transform8x8=1, slices=0)
    if arbiter.available():
        dec = arbiter.decode_annexb(g["stream"].tobytes())
        assert dec[0][0].shape == (768, 1366)
    # fast pan: 23 px / picture horizontally, 9 vertically (search range and border clamps)
    w, h = 320, 192
    base = rng.integers(0, 256, (h + 64, w + 256), dtype=np.uint8)
    base = (base.astype(np.int32) + np.roll(base, 1, 0) + np.roll(base, 1, 1) + np.roll(base, 2, 1)) // 4
    frames = []
    for i in range(6):
        y = base[9 * i % 60: 9 * i % 60 + h, 23 * i: 23 * i + w].astype(np.uint8)
        frames.append(np.concatenate([y.ravel(), np.full(w * h // 2, 128, np.uint8)]))
    pan = np.stack(frames)
    for ent in (0, 1):
        _same_as_oracle(w, h, pan, gop=6, qp_i=24, qp_p=26, entropy=ent, transform8x8=ent)
    # flat pictures: everything skips; black, white, mid-grey
    for v in (0, 128, 255):
        flat = np.full((4, w * h * 3 // 2), v, np.uint8)
        for ent in (0, 1):
            _same_as_oracle(w, h, flat, gop=4, qp_i=30, qp_p=30, entropy=ent)
    # bitrate target + High profile + slices + CABAC
    w, h = 640, 360
    _same_as_oracle(w, h, synth.make_clip(w, h, 45, seed=6), gop=20, fps=30, rc_mode=1, bitrate=1_500_000, entropy=1,
                    transform8x8=1, slices=3)
    # nv12 at 1080p through K1, High profile
    w, h = 1920, 1080
    yuv = synth.make_clip(w, h, 3, seed=9)
    nv = np.stack([np.concatenate([fr[: w * h], np.stack([fr[w * h: w * h + w * h // 4], fr[w * h + w * h // 4:]], -1).ravel()]) for fr in yuv])
    a = _same_as_oracle(w, h, nv, gop=3, qp_i=25, qp_p=27, in_fmt=1, entropy=1, transform8x8=1, slices=0)
    b = api.encode_frames(api.default_params(w, h, gop=3, qp_i=25, qp_p=27, entropy=1, transform8x8=1, slices=0), yuv)
    assert a["stream"].tobytes() == b["stream"].tobytes()
    # one session, two different clips back to back == two fresh sessions
    w, h = 320, 192
    c1, c2 = synth.make_clip(w, h, 9, seed=11), synth.make_clip(w, h, 7, seed=12)
    p = api.default_params(w, h, gop=4, qp_i=24, qp_p=26, entropy=1, transform8x8=1)
    with api.Session(p, 9) as ses:
        outs = []
        for c in (c1, c2, c1):
            ses.upload(c); ses.encode()
            outs.append(ses.download()["stream"].tobytes())
    assert outs[0] == outs[2] == api.encode_frames(p, c1)["stream"].tobytes()
    assert outs[1] == api.encode_frames(p, c2)["stream"].tobytes()


def test_one_task_sharded_across_gpus(built, tmp_path, monkeypatch):
    """VCPENC_GPUS=N: the closed GOPs of one task are encoded on N devices and concatenated on the host;
    the file must be byte-identical to the single-GPU one (needs >= 2 visible GPUs)."""
    if api.device_count() < 2:
        pytest.skip("needs two GPUs")
    w, h, n = 640, 360, 50
    clip = synth.make_clip(w, h, n, seed=7)
    y4m = tmp_path / "in.y4m"
    with open(y4m, "wb") as f:
        f.write(b"YUV4MPEG2 W%d H%d F30:1 Ip A1:1 C420jpeg\n" % (w, h))
        for fr in clip:
            f.write(b"FRAME\n" + fr.tobytes())
    outs = {}
    for gpus in ("1", "2", "all"):
        monkeypatch.setenv("VCPENC_GPUS", gpus)
        out = tmp_path / ("o%s.mp4" % gpus)
        api.transcode(str(y4m), str(out), "-c:v libx264 -crf 24 -g 8")
        outs[gpus] = out.read_bytes()
    api.thread_release()
    assert outs["1"] == outs["2"] == outs["all"]


def test_cabac_long_gops_and_bitrate_mode(built):
    """CABAC runs as batches of 8 pictures per GOP behind the reconstruction chain: GOPs longer than
    one batch, ragged last GOPs, many slices, and the bin-count fed rate control."""
    from oracle import pyoracle
    w, h = 176, 144
    for (n, gop, sl, kw) in ((21, 10, 1, {}), (19, 9, 3, {}), (30, 17, 2, dict(rc_mode=1, bitrate=600_000, fps=25))):
        clip = synth.make_clip(w, h, n, seed=40 + n)
        ref = pyoracle.encode(pyoracle.make_params(w, h, gop=gop, slices=sl, entropy=1, **kw), clip)
        got = api.encode_frames(api.default_params(w, h, gop=gop, slices=sl, entropy=1, **kw), clip, want_recon=True)
        assert [x[3] for x in got["info"]] == [x[3] for x in ref["info"]]
        assert got["stream"].tobytes() == ref["stream"], (n, gop, sl)
        assert np.array_equal(got["recon"], ref["recon"])
    # extreme content: white noise at low and high QP (long escape codes, outstanding-byte runs)
    noise = np.random.default_rng(1).integers(0, 256, (3, 64 * 64 * 3 // 2), dtype=np.uint8)
    for qp in (4, 51):
        ref = pyoracle.encode(pyoracle.make_params(64, 64, gop=60, qp_i=qp, qp_p=qp, entropy=1), noise)
        got = api.encode_frames(api.default_params(64, 64, gop=60, qp_i=qp, qp_p=qp, entropy=1), noise)
        assert got["stream"].tobytes() == ref["stream"], qp


def test_auto_slices_1080p_cabac(built):
    """slices=0 leaves the slice count to the encoder (vcp_auto_slices: 4 at 1080p with CABAC, 1 with CAVLC)."""
    from oracle import pyoracle
    w, h, n = 1920, 1080, 3
    clip = synth.make_clip(w, h, n, seed=31)
    for ent, want in ((1, 4), (0, 1)):
        ref = pyoracle.encode(pyoracle.make_params(w, h, gop=3, qp_i=25, qp_p=27, slices=0, entropy=ent), clip)
        got = api.encode_frames(api.default_params(w, h, gop=3, qp_i=25, qp_p=27, slices=0, entropy=ent), clip)
        assert got["stream"].tobytes() == ref["stream"]
        nal_types = [b[0] & 31 for b in got["stream"].tobytes().split(b"\x00\x00\x00\x01")[1:]]
        assert sum(t in (1, 5) for t in nal_types) == want * n


def test_intra_macroblocks_in_p_pictures(built):
    """A scene cut in the middle of a GOP: the refine flags macroblocks intra (estimate on the original
    picture), i_fix_kernel codes them on a wavefront after the inter ones.  Both entropy coders."""
    from oracle import pyoracle
    w, h = 320, 192
    a = synth.make_clip(w, h, 4, seed=1)
    cut = np.concatenate([a, np.roll(a, 7777, axis=1)[:2], synth.make_clip(w, h, 3, seed=77, start=9)])
    for kw in (dict(entropy=0), dict(entropy=1), dict(entropy=0, slices=3, deblock_idc=2), dict(entropy=1, slices=4)):
        ref = pyoracle.encode(pyoracle.make_params(w, h, gop=60, qp_i=26, qp_p=28, **kw), cut, want_dump=True)
        assert (ref["dump"]["mb_type"][4] == 0).sum() > 100       # the cut picture is mostly intra
        p = api.default_params(w, h, gop=60, qp_i=26, qp_p=28, debug=1, **kw)
        with api.Session(p, cut.shape[0]) as s:
            s.upload(cut)
            s.encode()
            got = s.download(want_recon=True)
            dbg = s.debug_mbs()
        assert np.array_equal(dbg["mb_type"], ref["dump"]["mb_type"]), kw
        assert np.array_equal(got["recon"], ref["recon"]), kw
        assert got["stream"].tobytes() == ref["stream"], kw
        if arbiter.available():
            dec = arbiter.decode_annexb(got["stream"].tobytes())
            for i in range(cut.shape[0]):
                assert np.array_equal(_flat(dec[i]), got["recon"][i])


def test_edge_cases(built):
    from oracle import pyoracle
    # smallest picture, one frame; GOP 1; one slice per macroblock row; ragged last GOP
    for (w, h, n, kw) in ((16, 16, 1, dict(gop=60)), (64, 64, 3, dict(gop=1)), (64, 64, 5, dict(gop=2, slices=4)),
                          (48, 80, 7, dict(gop=3, slices=5, deblock_idc=2))):
        clip = synth.make_clip(max(w, 32), max(h, 32), n, seed=3)[:, : ]
        if (w, h) == (16, 16):
            clip = np.random.default_rng(0).integers(0, 256, (n, 16 * 16 * 3 // 2), dtype=np.uint8)
        else:
            clip = synth.make_clip(w, h, n, seed=3)
        ref = pyoracle.encode(pyoracle.make_params(w, h, qp_i=22, qp_p=24, **kw), clip)
        got = api.encode_frames(api.default_params(w, h, qp_i=22, qp_p=24, **kw), clip, want_recon=True)
        assert got["stream"].tobytes() == ref["stream"], (w, h, kw)
        assert np.array_equal(got["recon"], ref["recon"])
    # static content collapses to P_Skip and still matches
    still = np.repeat(synth.make_clip(128, 96, 1, seed=8), 6, axis=0)
    ref = pyoracle.encode(pyoracle.make_params(128, 96, gop=60, qp_i=26, qp_p=28), still)
    got = api.encode_frames(api.default_params(128, 96, gop=60, qp_i=26, qp_p=28), still)
    assert got["stream"].tobytes() == ref["stream"]
    # extreme content: white noise at the lowest and highest QP
    noise = np.random.default_rng(1).integers(0, 256, (3, 64 * 64 * 3 // 2), dtype=np.uint8)
    for qp in (0, 51):
        ref = pyoracle.encode(pyoracle.make_params(64, 64, gop=60, qp_i=qp, qp_p=qp), noise)
        got = api.encode_frames(api.default_params(64, 64, gop=60, qp_i=qp, qp_p=qp), noise)
        assert got["stream"].tobytes() == ref["stream"], qp
    # invalid parameters surface as error classes, not crashes
    with pytest.raises(api.VcpencError) as e:
        api.encode_frames(api.default_params(64, 64, gop=0), noise)
    assert e.value.code == 1
    with pytest.raises(api.VcpencError) as e:
        api.encode_frames(api.default_params(64, 64, entropy=2), noise)
    assert e.value.code == 1


def test_hevc_intra_cus_in_p_pictures(built):
    """HEVC, a scene cut inside a GOP: the refine flags CUs intra, hevc_i_fix codes them on a wavefront after the
    inter ones; merge / AMVP skip them, deblocking filters their edges with bS 2 (chroma included)."""
    from oracle import pyoracle
    w, h = 320, 192
    a = synth.make_clip(w, h, 4, seed=1)
    cut = np.concatenate([a, np.roll(a, 7777, axis=1)[:2], synth.make_clip(w, h, 3, seed=77, start=9)])
    for kw in (dict(slices=1), dict(slices=3), dict(slices=2, deblock_idc=1)):
        ref = pyoracle.encode_hevc(pyoracle.make_params(w, h, codec=1, gop=60, qp_i=26, qp_p=28, **kw), cut)
        assert ref["info"][4][1] > 8 * ref["info"][3][1]          # the cut picture is mostly intra
        p = api.default_params(w, h, codec=1, gop=60, qp_i=26, qp_p=28, debug=1, **kw)
        with api.Session(p, cut.shape[0]) as s:
            s.upload(cut)
            s.encode()
            got = s.download(want_recon=True)
            dbg = s.debug_mbs()
        assert (dbg["mb_type"][4] == 0).sum() > 100, kw
        assert np.array_equal(got["recon"], ref["recon"]), kw
        assert got["stream"].tobytes() == ref["stream"], kw
        if arbiter.available():
            dec = arbiter.decode_annexb_hevc(got["stream"].tobytes())
            for i in range(cut.shape[0]):
                assert np.array_equal(_flat(dec[i]), got["recon"][i])


def test_hevc_half_sample_motion(built):
    """params.hevc_subpel: 8-tap half-sample planes per picture (k2_hpel.cu), half-sample step of the shared refine,
    plane fetch + four chroma phases in hevc_p_recon -- byte-identical to the oracle, on a clip that moves by half
    samples, on the standard clip and across a scene cut."""
    from oracle import pyoracle
    w, h, n = 320, 192, 6
    rng = np.random.default_rng(3)
    big = rng.integers(0, 256, (2 * h + 64, 2 * w + 64)).astype(np.float64)
    for _ in range(6):
        big = (big + np.roll(big, 1, 0) + np.roll(big, -1, 0) + np.roll(big, 1, 1) + np.roll(big, -1, 1)) / 5.0
    big = (big - big.min()) / (big.max() - big.min()) * 255.0
    half = np.stack([np.concatenate([np.clip(np.rint(big[i:i + 2 * h:2, 3 * i:3 * i + 2 * w:2]), 0, 255).astype(np.uint8).ravel(),
                                     np.full(w * h // 2, 128, np.uint8)]) for i in range(n)])
    a = synth.make_clip(w, h, 4, seed=1)
    cut = np.concatenate([a, np.roll(a, 7777, axis=1)[:2], synth.make_clip(w, h, 3, seed=77, start=9)])
    for clip, kw in ((half, dict(slices=1)), (half, dict(slices=3, deblock_idc=1)), (synth.make_clip(w, h, 8, seed=4), dict(slices=2)),
                     (cut, dict(slices=2)), (synth.make_clip(208, 114, 5, seed=9), dict(slices=1))):
        cw, chh = (208, 114) if clip.shape[1] == 208 * 114 * 3 // 2 else (w, h)
        gop = 60 if clip is cut else clip.shape[0]
        ref = pyoracle.encode_hevc(pyoracle.make_params(cw, chh, codec=1, gop=gop, qp_i=26, qp_p=28, hevc_subpel=1, **kw), clip)
        p = api.default_params(cw, chh, codec=1, gop=gop, qp_i=26, qp_p=28, hevc_subpel=1, debug=1, **kw)
        with api.Session(p, clip.shape[0]) as s:
            s.upload(clip)
            s.encode()
            got = s.download(want_recon=True)
            dbg = s.debug_mbs()
        bad = [i for i in range(clip.shape[0]) if not np.array_equal(got["recon"][i], ref["recon"][i])]
        assert not bad, ("recon differs", kw, bad)
        assert got["stream"].tobytes() == ref["stream"], kw
        if clip is half:
            assert (dbg["mv_final"] & 3).any()             # half-sample vectors were chosen


def test_hevc_quarter_sample_motion(built):
    """params.hevc_subpel = 2 / 3: the quarter-sample step of the refine (2: candidates ranked by averages of the half-sample
    planes, the shared H.264 code path; 3: by their exact prediction, me_refine_kernel<true>) and the motion compensation
    of quarter-sample vectors in hevc_p_recon (7/8-tap filters, vcp_hevc_qpel.cuh), exactly as the oracle and the decoder
    do: on clips that move by quarter samples in both directions, the hard clip (fractional pan + noise), the standard
    clip, a scene cut, a ragged size, SAO on, several slice counts; and it saves bits."""
    from oracle import pyoracle
    w, h, n = 320, 192, 6
    rng = np.random.default_rng(3)
    big = rng.integers(0, 256, (4 * h + 128, 4 * w + 128)).astype(np.float64)
    for _ in range(10):
        big = (big + np.roll(big, 1, 0) + np.roll(big, -1, 0) + np.roll(big, 1, 1) + np.roll(big, -1, 1)) / 5.0
    big = (big - big.min()) / (big.max() - big.min()) * 255.0

    def moving(dx, dy):   # (dx, dy) quarter samples per picture, either sign
        x0, y0 = (0 if dx >= 0 else -dx * n), (0 if dy >= 0 else -dy * n)
        return np.stack([np.concatenate([np.clip(np.rint(big[y0 + dy * i:y0 + dy * i + 4 * h:4, x0 + dx * i:x0 + dx * i + 4 * w:4]), 0, 255).astype(np.uint8).ravel(),
                                         np.full(w * h // 2, 128, np.uint8)]) for i in range(n)])
    a = synth.make_clip(w, h, 4, seed=1)
    cut = np.concatenate([a, np.roll(a, 7777, axis=1)[:2], synth.make_clip(w, h, 3, seed=77, start=9)])
    quarter = moving(5, 3)
    cases = ((quarter, dict(slices=1)), (moving(-3, 7), dict(slices=3, deblock_idc=1)), (moving(1, -1), dict(slices=2, hevc_sao=1)),
             (synth.make_hard_clip(w, h, 8, seed=6), dict(slices=2)), (synth.make_clip(w, h, 8, seed=4), dict(slices=2)),
             (cut, dict(slices=2)), (synth.make_clip(208, 114, 5, seed=9), dict(slices=1)))
    for sub, (clip, kw) in [(2, c) for c in cases] + [(3, c) for c in cases]:
        cw, chh = (208, 114) if clip.shape[1] == 208 * 114 * 3 // 2 else (w, h)
        gop = 60 if clip is cut else clip.shape[0]
        ref = pyoracle.encode_hevc(pyoracle.make_params(cw, chh, codec=1, gop=gop, qp_i=26, qp_p=28, hevc_subpel=sub, **kw), clip)
        p = api.default_params(cw, chh, codec=1, gop=gop, qp_i=26, qp_p=28, hevc_subpel=sub, debug=1, **kw)
        with api.Session(p, clip.shape[0]) as s:
            s.upload(clip)
            s.encode()
            got = s.download(want_recon=True)
            dbg = s.debug_mbs()
        bad = [i for i in range(clip.shape[0]) if not np.array_equal(got["recon"][i], ref["recon"][i])]
        assert not bad, ("recon differs", sub, kw, bad)
        assert got["stream"].tobytes() == ref["stream"], (sub, kw)
        if clip is quarter:
            assert (dbg["mv_final"] & 1).any()             # quarter-sample vectors were chosen
            halfonly = api.encode_frames(api.default_params(cw, chh, codec=1, gop=gop, qp_i=26, qp_p=28, hevc_subpel=1, **kw), clip)
            assert len(got["stream"]) < 0.9 * len(halfonly["stream"])
            if arbiter.available():
                dec = arbiter.decode_annexb_hevc(got["stream"].tobytes())
                for i in range(clip.shape[0]):
                    assert np.array_equal(np.concatenate([pl.ravel() for pl in dec[i]]), got["recon"][i])


def test_hevc_sample_adaptive_offset(built):
    """params.hevc_sao on the device (hevc_sao_kernel / hevc_sao_copy_kernel, SAO syntax in hevc_bins_kernel): identical
    to the oracle with and without deblocking / half-sample motion, several slice counts, a scene cut, a ragged size."""
    from oracle import pyoracle
    w, h = 320, 192
    a = synth.make_clip(w, h, 4, seed=1)
    cut = np.concatenate([a, np.roll(a, 7777, axis=1)[:2], synth.make_clip(w, h, 3, seed=77, start=9)])
    cases = ((synth.make_clip(w, h, 6, seed=4), w, h, dict(slices=1, hevc_subpel=1)),
             (synth.make_clip(w, h, 6, seed=4), w, h, dict(slices=3)),
             (synth.make_clip(w, h, 6, seed=5), w, h, dict(slices=2, deblock_idc=1, hevc_subpel=1)),
             (synth.make_clip(w, h, 5, seed=6), w, h, dict(slices=12)),
             (cut, w, h, dict(slices=2, hevc_subpel=1)),
             (synth.make_clip(208, 114, 5, seed=9), 208, 114, dict(slices=1, hevc_subpel=1)))
    for clip, cw, chh, kw in cases:
        gop = 60 if clip is cut else clip.shape[0]
        for qp in (27, 37):
            ref = pyoracle.encode_hevc(pyoracle.make_params(cw, chh, codec=1, gop=gop, qp_i=qp - 2, qp_p=qp, hevc_sao=1, **kw), clip)
            got = api.encode_frames(api.default_params(cw, chh, codec=1, gop=gop, qp_i=qp - 2, qp_p=qp, hevc_sao=1, debug=1, **kw), clip, want_recon=True)
            bad = [i for i in range(clip.shape[0]) if not np.array_equal(got["recon"][i], ref["recon"][i])]
            assert not bad, ("recon differs", kw, qp, bad)
            assert got["stream"].tobytes() == ref["stream"], (kw, qp)


def test_streamed_upload_is_identical(built):
    """upload(wait=False): the encode starts each GOP group when its frames have landed; same bytes as the
    synchronous upload, for both codecs, also when the session is reused and when GOPs do not fill the groups."""
    w, h = 320, 240
    for codec, n, gop in ((0, 40, 8), (1, 40, 8), (0, 13, 5), (1, 7, 60)):
        clip = synth.make_clip(w, h, n, seed=300 + n)
        p = api.default_params(w, h, codec=codec, gop=gop, qp_i=26, qp_p=28, slices=0)
        want = api.encode_frames(p, clip)["stream"].tobytes()
        with api.Session(p, n) as s:
            for _ in range(2):
                s.upload(clip, wait=False)
                s.encode()
                assert s.download()["stream"].tobytes() == want, (codec, n, gop)
            s.upload(clip)
            s.encode()
            assert s.download()["stream"].tobytes() == want


def test_transcode_hevc_presets(built, tmp_path):
    """The reference's h265-* presets (internal/config/config.go:47-50) through the transcode entry point: the MP4
    (hvc1 + hvcC) demuxes and decodes with FFmpeg to the pictures the raw .h265 output decodes to."""
    w, h, n = 640, 360, 20
    clip = synth.make_clip(w, h, n, seed=79)
    y4m = tmp_path / "in.y4m"
    with open(y4m, "wb") as f:
        f.write(b"YUV4MPEG2 W%d H%d F30:1 Ip A1:1 C420jpeg\n" % (w, h))
        for fr in clip:
            f.write(b"FRAME\n" + fr.tobytes())
    for name, args in (("cpu", "-c:v libx265 -preset medium -crf 28 -c:a aac -b:a 128k -movflags +faststart -g 10"),
                       ("nvenc", "-c:v hevc_nvenc -preset p4 -b:v 8M -c:a aac -b:a 128k -movflags +faststart -g 10"),
                       ("nvenc-hq", "-c:v hevc_nvenc -preset p7 -tune hq -b:v 10M -c:a aac -b:a 192k -movflags +faststart -g 10")):
        out = tmp_path / ("out_%s.mp4" % name)
        raw = tmp_path / ("out_%s.h265" % name)
        api.transcode(str(y4m), str(out), args)
        api.transcode(str(y4m), str(raw), args)
        api.verify(str(out))
        api.verify(str(raw))
        data = out.read_bytes()
        assert b"hvc1" in data and b"hvcC" in data and data.find(b"moov") < data.find(b"mdat")
        if arbiter.available():
            assert arbiter.probe_has_video(str(out))
            dec = arbiter.decode_file(str(out))
            ref = arbiter.decode_annexb_hevc(raw.read_bytes())
            assert len(dec) == n and len(ref) == n
            for i in range(n):
                assert np.array_equal(_flat(dec[i]), _flat(ref[i])), (name, i)
            assert arbiter.psnr(dec[5][0], synth.split_planes(clip[5], w, h)[0]) > 32


def test_transcode_drop_in(built, tmp_path):
    """The call the consumer makes: path in, path out, the reference's preset string, --verify."""
    w, h, n = 640, 360, 20
    clip = synth.make_clip(w, h, n, seed=77)
    y4m = tmp_path / "in.y4m"
    with open(y4m, "wb") as f:
        f.write(b"YUV4MPEG2 W%d H%d F30:1 Ip A1:1 C420jpeg\n" % (w, h))
        for fr in clip:
            f.write(b"FRAME\n" + fr.tobytes())
    out = tmp_path / "out.mp4"
    out.write_bytes(b"stale")                                     # -y: overwritten
    api.transcode(str(y4m), str(out), "-c:v libx264 -preset medium -crf 23 -c:a aac -b:a 128k -movflags +faststart -g 8")
    api.verify(str(out))
    data = out.read_bytes()
    assert data.find(b"moov") < data.find(b"mdat")
    # the calling thread keeps its session for the next task: same bytes from the reused session,
    # from a differently sized task in between, and after an explicit release
    api.transcode(str(y4m), str(out), "-c:v libx264 -preset medium -crf 23 -c:a aac -b:a 128k -movflags +faststart -g 8")
    assert out.read_bytes() == data
    api.thread_release()
    api.transcode(str(y4m), str(out), "-c:v libx264 -preset medium -crf 23 -c:a aac -b:a 128k -movflags +faststart -g 8")
    assert out.read_bytes() == data
    if arbiter.available():
        assert arbiter.probe_has_video(str(out))
        dec = arbiter.decode_file(str(out))
        assert len(dec) == n
        assert arbiter.psnr(dec[5][0], synth.split_planes(clip[5], w, h)[0]) > 32
    # argv door: exactly the reference's command line (cmd/consumer.go:376-380)
    libdir = os.path.dirname(api.LIB_PATH)
    out2 = tmp_path / "out2.mp4"
    r = subprocess.run([os.path.join(libdir, "vcp-ffmpeg"), "-hide_banner", "-loglevel", "warning", "-y", "-i", str(y4m)] +
                       "-c:v h264_nvenc -preset p4 -qp 28 -c:a aac -b:a 128k -movflags +faststart".split() + [str(out2)],
                       capture_output=True)
    assert r.returncode == 0, r.stderr
    pr = subprocess.run([os.path.join(libdir, "vcp-ffprobe"), "-v", "error", "-select_streams", "v:0", "-show_entries",
                         "stream=codec_type", "-of", "csv=p=0", str(out2)], capture_output=True)
    assert pr.returncode == 0 and b"video" in pr.stdout
    # K1 through the front door: a 4:4:4 y4m scaled down with -vf scale, and headerless rgb24 with -s
    y444 = tmp_path / "in444.y4m"
    with open(y444, "wb") as f:
        f.write(b"YUV4MPEG2 W%d H%d F30:1 Ip A1:1 C444\n" % (w, h))
        for fr in clip[:6]:
            y, u, v = synth.split_planes(fr, w, h)
            up = lambda c: np.repeat(np.repeat(c, 2, 0), 2, 1)
            f.write(b"FRAME\n" + y.tobytes() + up(u).tobytes() + up(v).tobytes())
    out4 = tmp_path / "out4.mp4"
    api.transcode(str(y444), str(out4), "-c:v libx264 -crf 23 -vf scale=320:-2 -g 3")
    api.verify(str(out4))
    if arbiter.available():
        dec = arbiter.decode_file(str(out4))
        assert len(dec) == 6 and dec[0][0].shape == (180, 320)
    rgb = tmp_path / "in.rgb"
    rgb.write_bytes(np.random.default_rng(2).integers(0, 256, 4 * 64 * 48 * 3, dtype=np.uint8).tobytes())
    out5 = tmp_path / "out5.h264"
    api.transcode(str(rgb), str(out5), "-c:v libx264 -qp 30 -s 64x48")
    if arbiter.available():
        assert len(arbiter.decode_annexb(out5.read_bytes())) == 4
    # container input (what the producer forwards): our own MP4 goes back in through the libavformat /
    # libavcodec front end and is re-encoded; -s after -i scales the output
    out6 = tmp_path / "out6.mp4"
    api.transcode(str(out), str(out6), "-c:v h264_nvenc -preset p4 -b:v 2M -g 10 -s 320x180")
    api.verify(str(out6))
    if arbiter.available():
        dec = arbiter.decode_file(str(out6))
        assert len(dec) == n and dec[0][0].shape == (180, 320)
        y5 = synth.split_planes(clip[5], w, h)[0].astype(np.int32)
        small = ((y5[0::2, 0::2] + y5[0::2, 1::2] + y5[1::2, 0::2] + y5[1::2, 1::2] + 2) >> 2).astype(np.uint8)
        assert arbiter.psnr(dec[5][0], small) > 24      # two generations of lossy coding at low bitrate
    # long inputs are processed in chunks of whole GOPs: forcing 2-GOP chunks must not change a byte
    import os as _os
    _os.environ["VCPENC_CHUNK_BYTES"] = str(16 * w * h * 3 // 2)
    try:
        out7 = tmp_path / "out7.mp4"
        api.transcode(str(y4m), str(out7), "-c:v libx264 -preset medium -crf 23 -c:a aac -b:a 128k -movflags +faststart -g 8")
        assert out7.read_bytes() == data
    finally:
        del _os.environ["VCPENC_CHUNK_BYTES"]
    # failure semantics: unknown container -> error class, no output left behind
    bad = tmp_path / "in.mkv"
    bad.write_bytes(b"\x1a\x45\xdf\xa3junk")
    out3 = tmp_path / "out3.mp4"
    with pytest.raises(api.VcpencError) as e:
        api.transcode(str(bad), str(out3), "-c:v libx264 -crf 23")
    assert e.value.code == 3 and not out3.exists()
    with pytest.raises(api.VcpencError) as e:
        api.transcode(str(y4m), str(out3), "-c copy")
    assert e.value.code == 8
    # cancellation flag (replaces exec.CommandContext's SIGKILL)
    import ctypes
    flag = ctypes.c_int(1)
    with pytest.raises(api.VcpencError) as e:
        api.transcode(str(y4m), str(out3), "-c:v libx264 -crf 23", cancel=flag)
    assert e.value.code == 6 and not out3.exists()


def test_effort_tier_fast_equals_oracle(built):
    """`-preset` fast tiers (effort 0): the refine stops at half samples; kernels and oracle agree byte for byte"""
    from oracle import pyoracle
    w, h, n = 320, 192, 7
    clip = synth.make_hard_clip(w, h, n, seed=5, noise=1)
    for entropy in (0, 1):
        ref = pyoracle.encode(pyoracle.make_params(w, h, gop=7, qp_i=24, qp_p=26, slices=2, effort=0, entropy=entropy), clip, want_dump=True)
        p = api.default_params(w, h, gop=7, qp_i=24, qp_p=26, slices=2, effort=0, entropy=entropy, debug=1)
        with api.Session(p, n) as s:
            s.upload(clip)
            s.encode()
            got = s.download(want_recon=True)
            dbg = s.debug_mbs()
        assert np.array_equal(dbg["mv_final"], ref["dump"]["mv_final"]) and not (dbg["mv_final"] & 1).any()
        assert got["stream"].tobytes() == ref["stream"] and np.array_equal(got["recon"], ref["recon"])
