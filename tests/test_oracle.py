"""The CPU oracle is pinned by the arbiter the north-star names: the FFmpeg h264 decoder must
decode the oracle's stream to exactly the oracle's reconstruction.  Plus hand-checkable
known answers and the committed golden hashes."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import CASES
from oracle import pyoracle
from video_codec_pipeline_b200 import arbiter, synth

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "h264_golden.json")))


def _lib():
    return pyoracle.lib()


def test_exp_golomb_known_answers():
    L = _lib()
    buf = (C.c_uint8 * 8)()
    # ue: 0->'1', 1->'010', 2->'011', 3->'00100', 7->'0001000'
    for k, bits in ((0, "1"), (1, "010"), (2, "011"), (3, "00100"), (7, "0001000"), (14, "0001111")):
        n = L.orc_ue_bits(k, buf)
        assert n == len(bits)
        got = "".join(format(b, "08b") for b in bytes(buf))[:n]
        assert got == bits
    # se: 0->'1', 1->'010', -1->'011', 2->'00100', -2->'00101'
    for v, bits in ((0, "1"), (1, "010"), (-1, "011"), (2, "00100"), (-2, "00101")):
        n = L.orc_se_bits(v, buf)
        got = "".join(format(b, "08b") for b in bytes(buf))[:n]
        assert got == bits


def test_transform_roundtrip_identity():
    """Forward core transform, then scaling equivalent to an ideal quantiser, then the decoder's
    inverse must give back the residual (Cf / Ci are inverses up to the known per-position norms)."""
    L = _lib()
    rng = np.random.default_rng(3)
    for _ in range(50):
        d = rng.integers(-255, 256, 16).astype(np.int32)
        w = np.zeros(16, np.int32)
        L.orc_fdct4(d.ctypes.data_as(C.POINTER(C.c_int)), w.ctypes.data_as(C.POINTER(C.c_int)))
        # norms: class 0 -> 1/16, class 1 -> 1/25, class 2 -> 1/20 ; inverse expects coefficients * 64
        cls = np.array([0, 2, 0, 2, 2, 1, 2, 1, 0, 2, 0, 2, 2, 1, 2, 1])
        scale = np.array([64 / 16.0, 64 / 25.0, 64 / 20.0])[cls]
        c = np.round(w * scale * 16).astype(np.int64)  # keep 4 extra bits of precision
        c = (c // 16).astype(np.int32)
        r = np.zeros(16, np.int32)
        L.orc_idct4(c.ctypes.data_as(C.POINTER(C.c_int)), r.ctypes.data_as(C.POINTER(C.c_int)))
        assert np.abs(r - d).max() <= 1


def test_dc_only_block():
    L = _lib()
    d = np.full(16, 10, np.int32)
    w = np.zeros(16, np.int32)
    L.orc_fdct4(d.ctypes.data_as(C.POINTER(C.c_int)), w.ctypes.data_as(C.POINTER(C.c_int)))
    assert w[0] == 160 and not w[1:].any()


def test_quant_dequant_tables():
    L = _lib()
    # QP 28 -> qp%6 = 4, qbits = 19: level = (|w| * 8192 + f) >> 19 for position (0,0)
    w = np.zeros(16, np.int32)
    w[0] = 6400
    lv = np.zeros(16, np.int16)
    L.orc_quant4x4(w.ctypes.data_as(C.POINTER(C.c_int)), 28, 0, 0, C.c_void_p(lv.ctypes.data))
    assert lv[0] == (6400 * 8192 + (1 << 19) // 6) >> 19
    c = np.zeros(16, np.int32)
    L.orc_dequant4x4(C.c_void_p(lv.ctypes.data), 28, 0, c.ctypes.data_as(C.POINTER(C.c_int)))
    assert c[0] == lv[0] * 16 << 4
    assert L.orc_lambda(12) == 1 and L.orc_lambda(24) == 4 and L.orc_lambda(36) == 16


def _bits(buf, n):
    return "".join(format(b, "08b") for b in bytes(buf))[:n]


def test_cavlc_block_known_answers():
    """Worked example of the standard's CAVLC description (widely reproduced): block
    0,3,-1,0 / 0,-1,1,0 / 1,0,0,0 / 0,0,0,0 in scan order [0,3,0,1,-1,-1,0,1,0...], nC=0
    -> 000010001110010111101101."""
    L = _lib()
    out = (C.c_uint8 * 64)()
    c = np.array([0, 3, 0, 1, -1, -1, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0], np.int16)
    n = L.orc_cavlc_block(C.c_void_p(c.ctypes.data), 16, 0, out, 64)
    assert _bits(out, n) == "000010001110010111101101"
    # all-zero block, nC=0 -> coeff_token '1'
    z = np.zeros(16, np.int16)
    n = L.orc_cavlc_block(C.c_void_p(z.ctypes.data), 16, 0, out, 64)
    assert _bits(out, n) == "1"
    # chroma DC, single +1 at position 0: coeff_token(T1=1,TC=1) '1', sign '0', total_zeros(0)='1'
    c4 = np.array([1, 0, 0, 0], np.int16)
    n = L.orc_cavlc_block(C.c_void_p(c4.ctypes.data), 4, -1, out, 64)
    assert _bits(out, n) == "101"


def test_luma_interpolation_flat_and_ramp():
    L = _lib()
    plane = np.full((32, 32), 77, np.uint8)
    for fy in range(4):
        for fx in range(4):
            assert L.orc_luma_qpel(C.c_void_p(plane.ctypes.data), 32, 10, 10, fx, fy) == 77
    ramp = np.tile((np.arange(32) * 4).astype(np.uint8), (32, 1))
    assert L.orc_luma_qpel(C.c_void_p(ramp.ctypes.data), 32, 10, 10, 2, 0) == 42   # half-way between 40 and 44
    assert L.orc_luma_qpel(C.c_void_p(ramp.ctypes.data), 32, 10, 10, 1, 0) == 41
    assert L.orc_luma_qpel(C.c_void_p(ramp.ctypes.data), 32, 10, 10, 0, 2) == 40


@pytest.mark.parametrize("entropy,t8", [(0, 0), (1, 0), (0, 1), (1, 1)], ids=["cavlc", "cabac", "cavlc-high", "cabac-high"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d_n%d_g%d_s%d_d%d_q%d" % c)
def test_oracle_stream_decodes_to_its_own_recon(case, entropy, t8):
    if not arbiter.available():
        pytest.skip("bundled FFmpeg decoder not present")
    w, h, n, gop, sl, idc, qp = case
    clip = synth.make_clip(w, h, n, seed=1000 + w + qp)
    p = pyoracle.make_params(w, h, gop=gop, qp_i=max(0, qp - 2), qp_p=qp, slices=sl, deblock_idc=idc, entropy=entropy,
                             transform8x8=t8)
    r = pyoracle.encode(p, clip)
    dec = arbiter.decode_annexb(r["stream"])
    assert len(dec) == n
    for i in range(n):
        flat = np.concatenate([pl.ravel() for pl in dec[i]])
        assert np.array_equal(flat, r["recon"][i]), "frame %d" % i
    # sanity: it is a real encode, not a passthrough
    assert len(r["stream"]) < clip.size // 2 or qp <= 12
    y = synth.split_planes(clip[n - 1], w, h)[0]
    assert arbiter.psnr(dec[n - 1][0], y) > (18 if qp > 45 else 28)


@pytest.mark.parametrize("g", GOLD, ids=lambda g: "%dx%d_q%d_s%d_e%d_t%d" % (g["w"], g["h"], g["qp"], g["slices"], g.get("entropy", 0), g.get("transform8x8", 0)))
def test_oracle_matches_golden(g):
    clip = synth.make_clip(g["w"], g["h"], g["frames"], seed=g["seed"])
    assert hashlib.sha256(clip.tobytes()).hexdigest() == g["clip_sha256"], "synthetic clip generator drifted"
    p = pyoracle.make_params(g["w"], g["h"], gop=g["gop"], qp_i=max(0, g["qp"] - 2), qp_p=g["qp"],
                             slices=g["slices"], deblock_idc=g["deblock_idc"], entropy=g.get("entropy", 0),
                             transform8x8=g.get("transform8x8", 0))
    r = pyoracle.encode(p, clip)
    assert [x[1] for x in r["info"]] == g["frame_sizes"]
    assert hashlib.sha256(r["stream"]).hexdigest() == g["stream_sha256"]
    assert hashlib.sha256(r["recon"].tobytes()).hexdigest() == g["recon_sha256"]


GOLD_HEVC = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hevc_golden.json")))


@pytest.mark.parametrize("g", GOLD_HEVC, ids=lambda g: "hevc_%dx%d_q%d_s%d_d%d_h%d" % (g["w"], g["h"], g["qp"], g["slices"], g["deblock_idc"], g["hevc_subpel"]))
def test_hevc_oracle_matches_golden(g):
    clip = synth.make_clip(g["w"], g["h"], g["frames"], seed=g["seed"])
    assert hashlib.sha256(clip.tobytes()).hexdigest() == g["clip_sha256"], "synthetic clip generator drifted"
    p = pyoracle.make_params(g["w"], g["h"], codec=1, gop=g["gop"], qp_i=max(0, g["qp"] - 2), qp_p=g["qp"],
                             slices=g["slices"], deblock_idc=g["deblock_idc"], hevc_subpel=g["hevc_subpel"])
    r = pyoracle.encode_hevc(p, clip)
    assert [x[1] for x in r["info"]] == g["frame_sizes"]
    assert hashlib.sha256(r["stream"]).hexdigest() == g["stream_sha256"]
    assert hashlib.sha256(r["recon"].tobytes()).hexdigest() == g["recon_sha256"]


def test_oracle_edge_cases():
    # single frame; GOP of 1 (all IDR); max slices = one per macroblock row
    w, h = 64, 64
    clip = synth.make_clip(w, h, 3, seed=5)
    for kw in (dict(gop=1), dict(gop=60, slices=4), dict(gop=2, slices=4, deblock_idc=2)):
        p = pyoracle.make_params(w, h, qp_i=20, qp_p=22, **kw)
        r = pyoracle.encode(p, clip)
        if arbiter.available():
            dec = arbiter.decode_annexb(r["stream"])
            assert len(dec) == 3
            for i in range(3):
                assert np.array_equal(np.concatenate([pl.ravel() for pl in dec[i]]), r["recon"][i])
    # a static clip must be almost entirely P_Skip
    still = np.repeat(clip[:1], 4, axis=0)
    r = pyoracle.encode(pyoracle.make_params(w, h, gop=60, qp_i=26, qp_p=28), still, want_dump=True)
    assert (r["dump"]["mb_type"][1:] == 2).mean() > 0.9
    # invalid parameters are rejected
    bad = pyoracle.make_params(w, h, gop=0)
    with pytest.raises(RuntimeError):
        pyoracle.encode(bad, clip)


def test_transform8x8_roundtrip_and_tables():
    """High profile 8x8: forward transform + quantiser + normative dequantiser + inverse transform give the
    residual back (flat at QP 0 within 1, the quantiser's MF table matches the forward transform's norms),
    and a DC-only block has the expected coefficient."""
    L = _lib()
    rng = np.random.default_rng(8)
    d = np.full(64, 5, np.int32)
    w = np.zeros(64, np.int32)
    L.orc_fdct8(d.ctypes.data_as(C.POINTER(C.c_int)), w.ctypes.data_as(C.POINTER(C.c_int)))
    assert w[0] == 5 * 64 and not w[1:].any()
    for qp, tol in ((0, 1), (12, 5), (24, 20)):   # quantiser step 0.625, 2.5, 10
        for _ in range(20):
            d = rng.integers(-200, 201, 64).astype(np.int32)
            r = np.zeros(64, np.int32)
            L.orc_roundtrip8(d.ctypes.data_as(C.POINTER(C.c_int)), qp, r.ctypes.data_as(C.POINTER(C.c_int)))
            assert np.abs(r - d).max() <= tol, (qp, np.abs(r - d).max())


def test_cabac_engine_known_answers():
    """The arithmetic coder restated from 9.3.4.2 against hand-checkable facts: context
    initialisation of ctxIdx 0 at QP 26 (m=20, n=-15 -> preCtxState 17 -> pStateIdx 46, valMPS 0),
    and CABAC streams are smaller than CAVLC ones while reconstructing the same pictures."""
    m, n = 20, -15
    pre = max(1, min(126, ((m * 26) >> 4) + n))
    assert pre == 17 and 63 - pre == 46
    w, h = 320, 180
    clip = synth.make_clip(w, h, 6, seed=9)
    a = pyoracle.encode(pyoracle.make_params(w, h, gop=3, qp_i=26, qp_p=28), clip)
    b = pyoracle.encode(pyoracle.make_params(w, h, gop=3, qp_i=26, qp_p=28, entropy=1), clip)
    assert np.array_equal(a["recon"], b["recon"])            # entropy coding is lossless
    assert len(b["stream"]) < 0.97 * len(a["stream"])


def test_oracle_intra_in_p_pictures():
    """Scene cut inside a GOP: most macroblocks of the cut picture turn Intra16x16, the stream stays
    decodable bit-exactly with both entropy coders, and quality recovers at once."""
    w, h = 320, 192
    a = synth.make_clip(w, h, 4, seed=1)
    cut = np.concatenate([a, np.roll(a, 7777, axis=1)[:3]])
    for ent in (0, 1):
        r = pyoracle.encode(pyoracle.make_params(w, h, gop=60, qp_i=26, qp_p=28, entropy=ent, slices=2), cut, want_dump=True)
        t = r["dump"]["mb_type"]
        assert (t[4] == 0).sum() > 100 and (t[1:4] == 0).sum() < 10
        if arbiter.available():
            dec = arbiter.decode_annexb(r["stream"])
            for i in range(cut.shape[0]):
                assert np.array_equal(np.concatenate([pl.ravel() for pl in dec[i]]), r["recon"][i])
            assert arbiter.psnr(dec[4][0], synth.split_planes(cut[4], w, h)[0]) > 35


def test_oracle_bitrate_mode():
    """-b:v: per-GOP budget, QP feedback two pictures late (vcp_algo.h); stream must stay decodable
    and land near the target."""
    w, h, n, fps = 320, 192, 48, 24
    clip = synth.make_clip(w, h, n, seed=12)
    sizes = {}
    for br in (400_000, 1_600_000):
        p = pyoracle.make_params(w, h, fps=fps, gop=24, rc_mode=1, bitrate=br)
        r = pyoracle.encode(p, clip)
        sizes[br] = len(r["stream"]) * 8 / (n / fps)
        qps = [x[3] for x in r["info"]]
        assert qps[0] == qps[1] - 3 and len(set(qps)) > 2          # IDR offset, feedback moves QP
        if arbiter.available():
            dec = arbiter.decode_annexb(r["stream"])
            for i in range(n):
                assert np.array_equal(np.concatenate([pl.ravel() for pl in dec[i]]), r["recon"][i])
    assert sizes[1_600_000] > 2.0 * sizes[400_000]
    assert 0.6 * 1_600_000 < sizes[1_600_000] < 1.5 * 1_600_000


def _vbv_min_fullness(bits, rate, buf, start):
    """Decoder buffer over the whole stream: +rate per picture interval, capped at buf, minus the picture."""
    f, lo = start, float("inf")
    for b in bits:
        f = min(buf, f + rate) - b
        lo = min(lo, f)
    return lo


def test_oracle_vbv_maxrate_bufsize():
    """-maxrate / -bufsize (the h264-nvenc-hq preset, /root/reference/internal/config/config.go:46): every GOP steers its own
    buffer model (vcp_algo.h: vcp_rc_vbv_qp); the concatenated stream must not underflow a buffer that starts at
    FFmpeg's 0.9 B, where the same clip at the plain constant QP does."""
    w, h, n, fps, gop = 320, 192, 72, 24, 24
    clip = synth.make_hard_clip(w, h, n, seed=5)
    for ent in (0, 1):
        kw = dict(fps=fps, gop=gop, qp_i=21, qp_p=24, entropy=ent)
        free = pyoracle.encode(pyoracle.make_params(w, h, **kw), clip)
        bits0 = [x[1] * 8 for x in free["info"]]
        natural = sum(bits0) / n * fps
        maxrate = int(0.7 * natural)
        bufsize = maxrate                                          # one second
        r = pyoracle.encode(pyoracle.make_params(w, h, maxrate=maxrate, bufsize=bufsize, **kw), clip)
        bits = [x[1] * 8 for x in r["info"]]
        qps = [x[3] for x in r["info"]]
        assert _vbv_min_fullness(bits0, maxrate / fps, bufsize, 0.9 * bufsize) < 0      # the cap binds on this clip
        assert _vbv_min_fullness(bits, maxrate / fps, bufsize, 0.9 * bufsize) >= 0
        assert sum(bits) / n * fps < 1.15 * maxrate
        assert qps[0] == 21 and qps[1] == 24 and max(qps) > 24 and min(q for i, q in enumerate(qps) if i % gop) >= 24
        assert qps[gop] == 21 and qps[gop + 1] == 24               # every GOP starts from the nominal QPs
        if arbiter.available():
            dec = arbiter.decode_annexb(r["stream"])
            for i in range(n):
                assert np.array_equal(np.concatenate([pl.ravel() for pl in dec[i]]), r["recon"][i])
        # a VBV needs both numbers (like libx264): -maxrate alone changes nothing in constant-QP mode
        alone = pyoracle.encode(pyoracle.make_params(w, h, maxrate=maxrate, **kw), clip[:gop])
        assert alone["stream"] == pyoracle.encode(pyoracle.make_params(w, h, **kw), clip[:gop])["stream"]
    # -b:v above -maxrate: the GOP budget follows -maxrate
    kw = dict(fps=fps, gop=gop, rc_mode=1, bitrate=900_000)
    capped = pyoracle.encode(pyoracle.make_params(w, h, maxrate=300_000, bufsize=600_000, **kw), clip)
    plain = pyoracle.encode(pyoracle.make_params(w, h, **kw), clip)
    rate = len(capped["stream"]) * 8 / (n / fps)
    assert rate < 1.3 * 300_000 < len(plain["stream"]) * 8 / (n / fps)
    # the h264-nvenc-hq numbers scaled to this picture size: -b:v 15M -maxrate 20M -bufsize 30M at 1080p30 never binds
    scale = (w * h) / (1920 * 1080)
    hq = dict(fps=30, gop=gop, rc_mode=1, bitrate=int(15e6 * scale))
    a = pyoracle.encode(pyoracle.make_params(w, h, maxrate=int(20e6 * scale), bufsize=int(30e6 * scale), **hq), clip)
    b = pyoracle.encode(pyoracle.make_params(w, h, **hq), clip)
    assert a["stream"] == b["stream"]
    bits = [x[1] * 8 for x in a["info"]]
    assert _vbv_min_fullness(bits, 20e6 * scale / 30, 30e6 * scale, 0.9 * 30e6 * scale) > 0


# ---- K1: other input formats and scaling (vcp_algo.h: vcp_rgb_*, vcp_scale_pos, vcp_bilerp) ----
def _k1_inputs(clip, w, h):
    """The same yuv420p content re-expressed in formats whose conversion back is lossless."""
    cw, ch = w // 2, h // 2
    out = {}
    nv12, p444, p422 = [], [], []
    for fr in clip:
        y, u, v = synth.split_planes(fr, w, h)
        nv12.append(np.concatenate([y.ravel(), np.stack([u, v], -1).ravel()]))
        up = lambda c: np.repeat(np.repeat(c, 2, 0), 2, 1)
        p444.append(np.concatenate([y.ravel(), up(u).ravel(), up(v).ravel()]))
        p422.append(np.concatenate([y.ravel(), np.repeat(u, 2, 0).ravel(), np.repeat(v, 2, 0).ravel()]))
    out[1] = np.stack(nv12); out[3] = np.stack(p444); out[4] = np.stack(p422)
    return out


def test_k1_formats_lossless_roundtrip():
    w, h, n = 96, 64, 3
    clip = synth.make_clip(w, h, n, seed=21)
    base = pyoracle.encode(pyoracle.make_params(w, h, gop=2), clip)
    for fmt, frames in _k1_inputs(clip, w, h).items():
        got = pyoracle.encode(pyoracle.make_params(w, h, gop=2, in_fmt=fmt), frames)
        assert got["stream"] == base["stream"], fmt


def test_k1_rgb_known_answers():
    """BT.601 limited range, 8-bit fixed point: grey -> Y = 16 + ((220 v + 128) >> 8), U = V = 128;
    primaries hit the textbook values; bgr24 is rgb24 with the channels swapped."""
    w, h = 32, 32
    p = pyoracle.make_params(w, h, gop=1, qp_i=8, qp_p=8, in_fmt=2)   # (QP<4 clamps |level| to the CAVLC escape range)
    for v in (0, 1, 127, 200, 255):
        rgb = np.full((1, w * h * 3), v, np.uint8)
        rec = pyoracle.encode(p, rgb)["recon"][0]
        y, u, vv = synth.split_planes(rec, w, h)
        assert abs(int(y[8, 8]) - (16 + ((220 * v + 128) >> 8))) <= 1 and abs(int(u[4, 4]) - 128) <= 1 and abs(int(vv[4, 4]) - 128) <= 1
    red = np.tile(np.array([255, 0, 0], np.uint8), w * h)[None]
    rec = pyoracle.encode(p, red)["recon"][0]
    y, u, vv = synth.split_planes(rec, w, h)
    assert abs(int(y[8, 8]) - 82) <= 1 and abs(int(u[4, 4]) - 90) <= 1 and abs(int(vv[4, 4]) - 240) <= 1
    rng = np.random.default_rng(5)
    rgb = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
    a = pyoracle.encode(pyoracle.make_params(w, h, gop=2, in_fmt=2), rgb.reshape(2, -1))
    b = pyoracle.encode(pyoracle.make_params(w, h, gop=2, in_fmt=5), rgb[..., ::-1].reshape(2, -1))
    assert a["stream"] == b["stream"]


def test_k1_scale():
    """Scaling: identity when the sizes agree, constant pictures stay constant, the decoder still
    reproduces the reconstruction, and a 2:1 reduction of a linear ramp stays a linear ramp."""
    w, h = 128, 96
    clip = synth.make_clip(w, h, 2, seed=4)
    a = pyoracle.encode(pyoracle.make_params(w, h, gop=2), clip)
    b = pyoracle.encode(pyoracle.make_params(w, h, gop=2, in_width=w, in_height=h), clip)
    assert a["stream"] == b["stream"]
    for (dw, dh) in ((64, 48), (208, 114), (96, 96)):
        p = pyoracle.make_params(dw, dh, gop=2, in_width=w, in_height=h)
        got = pyoracle.encode(p, clip)
        if arbiter.available():
            dec = arbiter.decode_annexb(got["stream"])
            for i in range(2):
                assert np.array_equal(np.concatenate([pl.ravel() for pl in dec[i]]), got["recon"][i])
    flat = np.full((1, w * h * 3 // 2), 77, np.uint8)
    rec = pyoracle.encode(pyoracle.make_params(48, 32, gop=1, qp_i=8, qp_p=8, in_width=w, in_height=h), flat)["recon"][0]
    assert np.abs(rec.astype(int) - 77).max() <= 1
    ramp = np.empty((1, w * h * 3 // 2), np.uint8)
    ramp[0, : w * h] = np.tile(np.arange(w, dtype=np.uint8) + 20, h)
    ramp[0, w * h:] = 128
    rec = pyoracle.encode(pyoracle.make_params(64, 48, gop=1, qp_i=8, qp_p=8, in_width=w, in_height=h), ramp)["recon"][0]
    y = rec[: 64 * 48].reshape(48, 64).astype(int)
    assert np.abs(y[10] - (2 * np.arange(64) + 20.5)).max() <= 1.5


# ---- HEVC oracle (oracle/hevc_oracle.inc.c): pinned by the FFmpeg hevc decoder -------------------------
@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d_n%d_g%d_s%d_d%d_q%d" % c)
def test_hevc_oracle_stream_decodes_to_its_own_recon(case):
    if not arbiter.available():
        pytest.skip("bundled FFmpeg decoder not present")
    w, h, n, gop, sl, idc, qp = case                       # idc 1: in-loop deblocking off
    clip = synth.make_clip(w, h, n, seed=1000 + w + qp)
    r = pyoracle.encode_hevc(pyoracle.make_params(w, h, codec=1, gop=gop, qp_i=max(0, qp - 2), qp_p=qp, slices=sl, deblock_idc=idc), clip)
    dec = arbiter.decode_annexb_hevc(r["stream"])
    assert len(dec) == n
    for i in range(n):
        assert np.array_equal(np.concatenate([pl.ravel() for pl in dec[i]]), r["recon"][i]), "frame %d" % i
    y = synth.split_planes(clip[n - 1], w, h)[0]
    assert arbiter.psnr(dec[n - 1][0], y) > (18 if qp > 45 else 28)
    assert len(r["stream"]) < clip.size // 2 or qp <= 12


def test_hevc_oracle_intra_cus_in_p_pictures():
    """Scene cut inside a GOP: CUs of the P picture are coded intra; the FFmpeg hevc decoder still agrees."""
    if not arbiter.available():
        pytest.skip("bundled FFmpeg decoder not present")
    w, h = 320, 192
    a = synth.make_clip(w, h, 4, seed=1)
    cut = np.concatenate([a, np.roll(a, 7777, axis=1)[:2], synth.make_clip(w, h, 3, seed=77, start=9)])
    for kw in (dict(slices=1), dict(slices=3), dict(slices=2, deblock_idc=1)):
        r = pyoracle.encode_hevc(pyoracle.make_params(w, h, codec=1, gop=60, qp_i=26, qp_p=28, **kw), cut)
        assert r["info"][4][1] > 8 * r["info"][3][1]
        dec = arbiter.decode_annexb_hevc(r["stream"])
        assert len(dec) == cut.shape[0]
        for i in range(cut.shape[0]):
            assert np.array_equal(np.concatenate([pl.ravel() for pl in dec[i]]), r["recon"][i]), (kw, i)
        assert arbiter.psnr(dec[4][0], synth.split_planes(cut[4], w, h)[0]) > 34


def test_hevc_oracle_half_sample_motion():
    """params.hevc_subpel (oracle only this round): 8-tap luma interpolation + the full set of 4-tap chroma phases.
    The decoder must agree bit-exactly, and on a clip that moves by half samples the tool must pay off."""
    if not arbiter.available():
        pytest.skip("bundled FFmpeg decoder not present")
    w, h, n = 320, 192, 6
    rng = np.random.default_rng(3)
    big = rng.integers(0, 256, (2 * h + 64, 2 * w + 64)).astype(np.float64)
    for _ in range(6):                                    # cheap low-pass: box blur, separable
        big = (big + np.roll(big, 1, 0) + np.roll(big, -1, 0) + np.roll(big, 1, 1) + np.roll(big, -1, 1)) / 5.0
    big = (big - big.min()) / (big.max() - big.min()) * 255.0
    clip = np.stack([np.concatenate([np.clip(np.rint(big[i:i + 2 * h:2, 3 * i:3 * i + 2 * w:2]), 0, 255).astype(np.uint8).ravel(),
                                     np.full(w * h // 2, 128, np.uint8)]) for i in range(n)])
    size = {}
    for sub in (0, 1):
        for kw in (dict(slices=1), dict(slices=3, deblock_idc=1)):
            r = pyoracle.encode_hevc(pyoracle.make_params(w, h, codec=1, gop=n, qp_i=26, qp_p=28, hevc_subpel=sub, **kw), clip)
            dec = arbiter.decode_annexb_hevc(r["stream"])
            assert len(dec) == n
            for i in range(n):
                assert np.array_equal(np.concatenate([pl.ravel() for pl in dec[i]]), r["recon"][i]), (sub, kw, i)
            size[(sub, kw["slices"])] = len(r["stream"])
    assert size[(1, 1)] < 0.7 * size[(0, 1)] and size[(1, 3)] < 0.7 * size[(0, 3)]
    # the standard clip (integer motion, noise patch, scene content): still exact
    clip2 = synth.make_clip(w, h, 6, seed=4)
    r = pyoracle.encode_hevc(pyoracle.make_params(w, h, codec=1, gop=6, qp_i=26, qp_p=28, hevc_subpel=1, slices=2), clip2)
    dec = arbiter.decode_annexb_hevc(r["stream"])
    for i in range(6):
        assert np.array_equal(np.concatenate([pl.ravel() for pl in dec[i]]), r["recon"][i])


def test_hevc_oracle_quarter_sample_motion():
    """params.hevc_subpel = 2 / 3: the 7/8-tap quarter-sample filters (candidates ranked by the half-sample proxy / by the
    exact prediction).  Decoder-exact, each precision step pays off on a clip that moves by quarter samples, and the
    proxy ranking finds what the exact ranking finds."""
    if not arbiter.available():
        pytest.skip("bundled FFmpeg decoder not present")
    w, h, n = 320, 192, 5
    rng = np.random.default_rng(3)
    big = rng.integers(0, 256, (4 * h + 64, 4 * w + 64)).astype(np.float64)
    for _ in range(12):
        big = (big + np.roll(big, 1, 0) + np.roll(big, -1, 0) + np.roll(big, 1, 1) + np.roll(big, -1, 1)) / 5.0
    big = (big - big.min()) / (big.max() - big.min()) * 255.0
    clip = np.stack([np.concatenate([np.clip(np.rint(big[i:i + 4 * h:4, 3 * i:3 * i + 4 * w:4]), 0, 255).astype(np.uint8).ravel(),
                                     np.full(w * h // 2, 128, np.uint8)]) for i in range(n)])
    size = {}
    for sub in (0, 1, 2, 3):
        r = pyoracle.encode_hevc(pyoracle.make_params(w, h, codec=1, gop=n, qp_i=26, qp_p=28, hevc_subpel=sub, slices=2), clip)
        dec = arbiter.decode_annexb_hevc(r["stream"])
        assert len(dec) == n
        for i in range(n):
            assert np.array_equal(np.concatenate([pl.ravel() for pl in dec[i]]), r["recon"][i]), (sub, i)
        size[sub] = len(r["stream"])
    assert size[2] < size[1] < size[0]
    assert abs(size[2] - size[3]) < 0.01 * size[3]


def test_hevc_oracle_sample_adaptive_offset():
    """params.hevc_sao: luma edge offsets decided per coding tree block on the deblocked picture.  The decoder must
    reproduce the oracle's reconstruction (syntax + process), with and without deblocking, 1..12 slices; SAO must
    not lower the PSNR it optimises."""
    if not arbiter.available():
        pytest.skip("bundled FFmpeg decoder not present")
    w, h, n = 320, 192, 6
    for seed, kw in ((4, dict(slices=1)), (4, dict(slices=3)), (5, dict(slices=2, deblock_idc=1)), (6, dict(slices=12))):
        clip = synth.make_clip(w, h, n, seed=seed)
        ps = {}
        for sao in (0, 1):
            r = pyoracle.encode_hevc(pyoracle.make_params(w, h, codec=1, gop=n, qp_i=25, qp_p=27, hevc_sao=sao, hevc_subpel=1, **kw), clip)
            dec = arbiter.decode_annexb_hevc(r["stream"])
            assert len(dec) == n
            for i in range(n):
                assert np.array_equal(np.concatenate([pl.ravel() for pl in dec[i]]), r["recon"][i]), (kw, sao, i)
            ps[sao] = np.mean([arbiter.psnr(dec[i][0], synth.split_planes(clip[i], w, h)[0]) for i in range(n)])
        assert ps[1] >= ps[0] - 0.01, (kw, ps)


def test_hevc_oracle_intra_modes():
    """params.hevc_intra_modes (oracle only this round): planar / DC / horizontal / vertical per intra CU, signalled
    through the MPM list or rem_intra_luma_pred_mode, with the mode-dependent coefficient scans (horizontal modes
    scan vertically and vice versa) and the smoothed references of 8x8 planar.  The decoder must agree bit-exactly
    in IDR pictures, across a scene cut inside a GOP, with SAO / half-sample motion on, and at a ragged size."""
    if not arbiter.available():
        pytest.skip("bundled FFmpeg decoder not present")
    w, h, n = 320, 192, 6
    a = synth.make_clip(w, h, 4, seed=1)
    cut = np.concatenate([a, np.roll(a, 7777, axis=1)[:2], synth.make_clip(w, h, 3, seed=77, start=9)])
    total = {0: 0, 1: 0}
    for clip, cw, chh, gop, kw in ((synth.make_clip(w, h, n, seed=4), w, h, 2, dict(slices=3)),
                                   (synth.make_clip(w, h, n, seed=5), w, h, 6, dict(slices=2, deblock_idc=1, hevc_sao=1)),
                                   (cut, w, h, 60, dict(slices=2, hevc_subpel=1, hevc_sao=1)),
                                   (synth.make_clip(208, 114, 5, seed=9), 208, 114, 5, dict(slices=1))):
        for qp in (22, 32):
            for im in (0, 1):
                r = pyoracle.encode_hevc(pyoracle.make_params(cw, chh, codec=1, gop=gop, qp_i=qp - 2, qp_p=qp, hevc_intra_modes=im, **kw), clip)
                total[im] += len(r["stream"])
                if im:
                    dec = arbiter.decode_annexb_hevc(r["stream"])
                    assert len(dec) == clip.shape[0]
                    for i in range(clip.shape[0]):
                        assert np.array_equal(np.concatenate([pl.ravel() for pl in dec[i]]), r["recon"][i]), (kw, qp, i)
    assert total[1] < total[0]                             # the extra modes pay for their signalling


def test_hevc_tables_match_decoder_rodata():
    """CABAC initValues (tables 9-5..9-37) typed in the oracle must appear in the decoder's own tables."""
    import glob
    import re
    d = arbiter._find_libdir()
    if d is None:
        pytest.skip("bundled libavcodec not present")
    data = open(sorted(glob.glob(os.path.join(d, "libavcodec-*.so*")))[0], "rb").read()
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "video_codec_pipeline_b200", "csrc", "hevc_tables.h")).read()
    body = re.search(r"hevc_init_values\[2\]\[HC_NCTX\] = \{(.*?)\}\};", src, re.S).group(1)
    rows = [[int(x) for x in re.findall(r"\b\d+\b", re.sub(r"/\*.*?\*/", "", part))] for part in body.split("},")]
    assert [len(r) for r in rows] == [132, 132]
    assert [r[130:] for r in rows] == [[153, 200], [153, 185]]      # sao_merge_flag, sao_type_idx (tables 9-5, 9-6)
    for row in rows:
        # last_sig_coeff prefix (18), coded_sub_block (4), sig_coeff (42), greater1 (24), greater2 (6) are contiguous runs
        for a, b in ((18, 36), (54, 58), (58, 100), (100, 124), (124, 130)):
            assert data.find(bytes(row[a:b])) >= 0, (a, b)


def test_oracle_coefficient_decimation():
    """vcp_algo.h: vcp_decimate_score / vcp_decimate_zero (inter macroblocks, medium and slow tiers): known answers of the
    score, and on a noisy clip the rule saves bits for a small PSNR cost while the stream stays decoder-exact."""
    import ctypes as C
    L = pyoracle.lib()
    L.orc_decimate_score.argtypes = [C.c_ulonglong, C.c_int, C.c_int]
    sc = lambda mask, big=0, is8=0: L.orc_decimate_score(mask, big, is8)
    assert sc(0) == 0 and sc(1) == 3 and sc(1 << 15) == 0 and sc(0b11) == 6 and sc(1 << 2) == 2 and sc(1 << 5) == 1
    assert sc(1, 1) == 9 and sc(0xffff) == 9                         # a level beyond +-1 / many levels: never decimated
    assert sc(1, 0, 1) == 3 and sc(1 << 40, 0, 1) == 0 and sc(1 << 5, 0, 1) == 2 and sc(1 << 20, 0, 1) == 1
    assert sc((1 << 3) | (1 << 40), 0, 1) == 3                       # run of 3 -> 3, then a run of 36 -> 0
    w, h, n = 320, 192, 8
    clip = synth.make_hard_clip(w, h, n, seed=9)
    for kw in (dict(entropy=0), dict(entropy=1, transform8x8=1)):
        on = pyoracle.encode(pyoracle.make_params(w, h, gop=n, qp_i=28, qp_p=30, effort=1, **kw), clip)
        off = pyoracle.encode(pyoracle.make_params(w, h, gop=n, qp_i=28, qp_p=30, effort=0, hevc_subpel=0, **kw), clip)
        # effort 0 also drops the quarter-sample step, which costs bits: the decimated stream must be clearly smaller
        assert len(on["stream"]) < 0.97 * len(off["stream"]), kw
        if arbiter.available():
            dec = arbiter.decode_annexb(on["stream"])
            assert all(np.array_equal(np.concatenate([pl.ravel() for pl in dec[i]]), on["recon"][i]) for i in range(n))
            ysz = w * h
            p_on = np.mean([arbiter.psnr(on["recon"][i][:ysz].reshape(h, w), clip[i][:ysz].reshape(h, w)) for i in range(1, n)])
            p_off = np.mean([arbiter.psnr(off["recon"][i][:ysz].reshape(h, w), clip[i][:ysz].reshape(h, w)) for i in range(1, n)])
            assert p_on > p_off - 0.6, (kw, p_on, p_off)


def test_oracle_effort_tiers():
    """`-preset` tiers (vcpenc_params.effort): the fast tiers stop the motion refine at half samples.  On a clip that
    pans by quarter samples the medium tier needs clearly fewer bits; both streams decode to their reconstruction."""
    w, h, n = 176, 144, 6
    clip = synth.make_hard_clip(w, h, n, seed=5, noise=0)          # (1.25, 0.75) samples per frame
    sizes = {}
    for effort in (0, 1, 2):
        r = pyoracle.encode(pyoracle.make_params(w, h, gop=6, qp_i=24, qp_p=26, effort=effort), clip, want_dump=True)
        sizes[effort] = len(r["stream"])
        mv = r["dump"]["mv_final"][1:]
        if effort == 0:
            assert not (mv & 1).any(), "fast tier: vectors stay on half-sample positions"
        else:
            assert (mv & 1).any()
        if arbiter.available():
            dec = arbiter.decode_annexb(r["stream"])
            assert all(np.array_equal(np.concatenate([pl.ravel() for pl in dec[i]]), r["recon"][i]) for i in range(n))
    assert sizes[1] == sizes[2] and sizes[1] < 0.985 * sizes[0], sizes   # IDR included: the P pictures alone differ more
