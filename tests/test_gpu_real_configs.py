"""GPU parity at the configurations the benchmarks and the built-in presets actually run (BASELINE.json configs[1..3],
/root/reference/internal/config/config.go:45-50): full picture sizes with the real GOP length of 60, at least two GOPs
(so the GOP boundary, idr_pic_id alternation and the lock-step batch are exercised), the presets as vcpenc_parse_args
sees them (High profile, CABAC), HEVC, and content on which no macroblock takes a cheap path.
Bar: the CUDA bitstream is byte-identical to the CPU oracle's and the FFmpeg decoder reproduces the encoder's
reconstruction exactly.  The scalar oracle needs ~13 s per 1080p GOP and ~55 s per 4K GOP; GOPs run on separate threads."""
import sys
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import media_util as mu  # noqa: E402

from video_codec_pipeline_b200 import api, arbiter, synth  # noqa: E402

pytestmark = pytest.mark.gpu

H264_CPU = "-c:v libx264 -preset medium -crf 23 -c:a aac -b:a 128k -movflags +faststart"          # config.go:49
H264_NVENC_HQ = "-c:v h264_nvenc -preset p7 -tune hq -b:v 15M -maxrate 20M -bufsize 30M -c:a aac -b:a 192k -movflags +faststart"  # config.go:46
H265_NVENC = "-c:v hevc_nvenc -preset p4 -b:v 8M -c:a aac -b:a 128k -movflags +faststart"          # config.go:47


def _flat(planes):
    return np.concatenate([pl.ravel() for pl in planes])


def _oracle_gops(w, h, clip, gop, kw, codec):
    """one oracle encode per closed GOP, in parallel; first_gop keeps idr_pic_id alternating like the unsharded encode"""
    from oracle import pyoracle
    n = clip.shape[0]
    ngop = (n + gop - 1) // gop

    def one(g):
        p = pyoracle.make_params(w, h, gop=gop, first_gop=g, **kw)
        fr = clip[g * gop:(g + 1) * gop]
        return pyoracle.encode_hevc(p, fr) if codec else pyoracle.encode(p, fr, want_recon=True)

    pyoracle.lib()
    with ThreadPoolExecutor(ngop) as ex:
        return list(ex.map(one, range(ngop)))


def _check(w, h, clip, gop, kw, decode_all=True):
    codec = kw.get("codec", 0)
    n = clip.shape[0]
    p = api.default_params(w, h, gop=gop, debug=1, **kw)
    with api.Session(p, n) as s:
        s.upload(clip, wait=False)
        s.encode()
        got = s.download(want_recon=True)
    refs = _oracle_gops(w, h, clip, gop, kw, codec)
    stream = b"".join(r["stream"] for r in refs)
    recon = np.concatenate([r["recon"] for r in refs], axis=0)
    assert got["stream"].tobytes() == stream, "CUDA bitstream differs from the oracle"
    assert np.array_equal(got["recon"], recon), "CUDA reconstruction differs from the oracle"
    if arbiter.available():
        data = got["stream"].tobytes()
        dec = arbiter.decode_annexb_hevc(data) if codec else arbiter.decode_annexb(data)
        assert len(dec) == n
        idx = range(n) if decode_all else (0, gop - 1, gop, n - 1)
        for i in idx:
            assert np.array_equal(_flat(dec[i]), got["recon"][i]), "decoder output != encoder reconstruction (frame %d)" % i
    return got


@pytest.mark.timeout(600)
def test_config1_1080p_gop60_cavlc(built):
    """BASELINE.json configs[1] as bench.py times it: 1080p30, GOP 60, CAVLC, 1 slice, CQP 25/27 -- two whole GOPs"""
    w, h, gop = 1920, 1080, 60
    clip = np.concatenate([synth.make_clip(w, h, gop, seed=1080, start=0), synth.make_clip(w, h, gop, seed=1081, start=gop)])
    _check(w, h, clip, gop, dict(qp_i=25, qp_p=27, slices=1, entropy=0))


@pytest.mark.timeout(600)
def test_h264_cpu_preset_as_parsed_1080p_gop60(built):
    """what `h264-cpu` (config.go:49) parses to -- High profile, CABAC, 8x8 transform, encoder-chosen slices, CRF 23 -> QP 21/24"""
    pp = api.parse_args(H264_CPU.split())
    assert pp.entropy == 1 and pp.transform8x8 == 1 and pp.gop == 60
    w, h, gop = 1920, 1080, 60
    clip = np.concatenate([synth.make_clip(w, h, gop, seed=1080, start=0), synth.make_clip(w, h, gop, seed=1081, start=gop)])
    _check(w, h, clip, gop, dict(qp_i=pp.qp_i, qp_p=pp.qp_p, slices=4, entropy=1, transform8x8=1))


@pytest.mark.timeout(600)
def test_hard_content_1080p_gop60_high_cabac(built):
    """full-frame fractional pan + noise: every macroblock searches sub-sample positions, codes a residual and feeds the
    arithmetic coder; one whole GOP + the start of a second one"""
    w, h, gop = 1920, 1080, 60
    clip = synth.make_hard_clip(w, h, gop + 6, seed=7)
    got = _check(w, h, clip, gop, dict(qp_i=21, qp_p=24, slices=4, entropy=1, transform8x8=1), decode_all=False)
    assert got["stream"].size > 3 * 1000 * 1000      # not a skip-dominated clip


@pytest.mark.timeout(600)
def test_hevc_1080p_gop60(built):
    """the h265-* presets' path at 1080p, GOP 60, two GOPs"""
    w, h, gop = 1920, 1080, 60
    clip = np.concatenate([synth.make_clip(w, h, gop, seed=1080, start=0), synth.make_clip(w, h, gop, seed=1081, start=gop)])
    _check(w, h, clip, gop, dict(codec=1, qp_i=26, qp_p=29, slices=4, hevc_subpel=2))


@pytest.mark.timeout(900)
def test_config2_4k_high_gop60(built):
    """BASELINE.json configs[2]: 4K60 High profile CABAC, GOP 60, 7 slices -- one whole GOP and the first pictures of the next"""
    w, h, gop = 3840, 2160, 60
    clip = synth.make_clip(w, h, gop + 4, seed=2160)
    _check(w, h, clip, gop, dict(qp_i=25, qp_p=27, slices=7, entropy=1, transform8x8=1, fps=60), decode_all=False)


# ---- the plugin call on real-looking inputs: foreign containers, foreign codecs, audio -------------------------------
def _source_avi(tmp_path, w=640, h=360, n=24, fps=24, rate=48000, channels=2):
    clip = synth.make_clip(w, h, n, seed=5)
    pcm = mu.tone(rate, n / fps, channels)
    path = str(tmp_path / "in.avi")
    mu.write_avi(path, clip, w, h, fps, pcm, rate)
    return path, clip, pcm


@pytest.mark.timeout(300)
def test_transcode_input_with_audio_h264_cpu_preset(built, tmp_path):
    """An input with an audio track through the VERBATIM h264-cpu preset string (it carries `-c:a aac -b:a 128k`): the
    output plays video and audio.  Video: FFmpeg's decode of the MP4 equals the oracle's reconstruction for the parsed
    parameters.  Audio: the AAC track decodes back to the source signal (AAC is lossy: SNR > 20 dB, alignment exact)."""
    from oracle import pyoracle
    src, clip, pcm = _source_avi(tmp_path)
    out = str(tmp_path / "out.mp4")
    api.transcode(src, out, H264_CPU + " -g 12")
    api.verify(out)
    raw = open(out, "rb").read()
    assert raw.find(b"moov") < raw.find(b"mdat") and b"mp4a" in raw and b"avc1" in raw
    pp = api.parse_args((H264_CPU + " -g 12").split())
    ref = pyoracle.encode(pyoracle.make_params(640, 360, fps=24, gop=12, qp_i=pp.qp_i, qp_p=pp.qp_p, entropy=1, transform8x8=1,
                                               slices=max(1, ((360 + 15) // 16) // 17)), clip, want_recon=True)
    dec = arbiter.decode_file(out)
    assert len(dec) == clip.shape[0]
    assert all(np.array_equal(_flat(dec[i]), ref["recon"][i]) for i in range(len(dec)))
    cid, au = arbiter.decode_audio_file(out)
    assert cid == 0x15002 and au.shape[1] == 2
    r = pcm.astype(np.float32) / 32768
    for c in range(2):
        lag, snr = mu.best_lag_snr(r[:, c], au[:, c], 3000)
        assert lag == 0 and snr > 20.0, (lag, snr)


@pytest.mark.timeout(300)
def test_transcode_aac_input_is_copied_and_presets_with_audio(built, tmp_path):
    """MP4 + AAC in (our own output) -> the bitrate presets: audio access units are stream-copied, video re-encoded"""
    src, clip, _ = _source_avi(tmp_path)
    first = str(tmp_path / "first.mp4")
    api.transcode(src, first, H264_CPU + " -g 12")
    a0 = api.probe_audio(first)
    assert a0["copied"]
    for k, preset in enumerate((H264_NVENC_HQ, H265_NVENC)):
        out = str(tmp_path / ("second%d.mp4" % k))
        api.transcode(first, out, preset + " -g 12")
        api.verify(out)
        a1 = api.probe_audio(out)
        assert a1["frames"] == a0["frames"] and a1["asc"] == a0["asc"]
        dec = arbiter.decode_file(out)
        assert len(dec) == clip.shape[0]
        assert arbiter.psnr(_flat(dec[3]), clip[3]) > 28.0


@pytest.mark.timeout(300)
def test_transcode_foreign_containers_and_codecs(built, tmp_path):
    """what the producer forwards (.mkv .avi .mov .webm, cmd/producer.go:485-488) with codecs that are not ours:
    MPEG-4 part 2, VP9, MJPEG written by OpenCV's FFmpeg back end"""
    cv2 = pytest.importorskip("cv2")
    w, h, n = 320, 240, 18
    made = []
    for name, fourcc in (("a.avi", "mp4v"), ("b.mkv", "mp4v"), ("c.webm", "VP90"), ("e.mov", "mp4v"), ("f.avi", "MJPG")):
        path = str(tmp_path / name)
        vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*fourcc), 30, (w, h))
        if not vw.isOpened():
            continue
        for i in range(n):
            img = np.zeros((h, w, 3), np.uint8)
            img[:, :, 0] = (np.arange(w) + 3 * i) % 256
            img[:, :, 1] = 40 + 5 * i
            img[40:120, 20 + 4 * i:100 + 4 * i, 2] = 230
            vw.write(img)
        vw.release()
        if os.path.getsize(path) > 0:
            made.append(path)
    assert len(made) >= 3
    for path in made:
        info = api.probe_input(path, max_frames=n)
        out = path + ".out.mp4"
        api.transcode(path, out, H264_CPU + " -g 9")
        api.verify(out)
        dec = arbiter.decode_file(out)
        assert len(dec) == info["frames"].shape[0] == n, path
        assert arbiter.psnr(_flat(dec[n // 2]), info["frames"][n // 2]) > 30.0, path


@pytest.mark.timeout(300)
def test_transcode_uploads_while_reading(built, tmp_path, monkeypatch):
    """raw / y4m inputs: the copies of a chunk are queued while the reader still delivers it (vcpenc_session_upload_gated).
    Same file as with the path switched off; a y4m whose FRAME markers carry parameters (the picture count computed from
    the file size is then too high: the chunk comes up short) and a file cut in the middle of a picture fall back to the
    plain upload and still give the right pictures; headerless raw input (.yuv, -s WxH) takes the same path."""
    w, h, n = 320, 192, 47
    clip = synth.make_clip(w, h, n, seed=33)
    hdr = b"YUV4MPEG2 W%d H%d F30:1 Ip A1:1 C420\n" % (w, h)
    bare = str(tmp_path / "bare.y4m")
    with open(bare, "wb") as f:
        f.write(hdr)
        for fr in clip:
            f.write(b"FRAME\n" + fr.tobytes())
    marked = str(tmp_path / "marked.y4m")
    with open(marked, "wb") as f:
        f.write(hdr)
        for fr in clip:
            f.write(b"FRAME Ip\n" + fr.tobytes())
    cut = str(tmp_path / "cut.y4m")
    with open(cut, "wb") as f:
        f.write(open(bare, "rb").read()[:len(hdr) + 30 * (6 + clip.shape[1]) + 1000])      # 30 whole pictures + a stump
    raw = str(tmp_path / "in.yuv")
    clip.tofile(raw)
    args = H264_CPU + " -g 6"
    outs = {}
    for chunk in (None, 6 * 2 * clip.shape[1]):          # one chunk; chunks of two GOPs
        if chunk:
            monkeypatch.setenv("VCPENC_CHUNK_BYTES", str(chunk))
        for name, src, a in (("bare", bare, args), ("marked", marked, args), ("raw", raw, "-s %dx%d -r 30 " % (w, h) + args)):
            out = str(tmp_path / ("%s_%s.mp4" % (name, chunk)))
            api.transcode(src, out, a)
            outs[(name, chunk)] = open(out, "rb").read()
        monkeypatch.setenv("VCPENC_NO_EARLY_UPLOAD", "1")
        off = str(tmp_path / ("off_%s.mp4" % chunk))
        api.transcode(bare, off, args)
        monkeypatch.delenv("VCPENC_NO_EARLY_UPLOAD")
        assert outs[("bare", chunk)] == open(off, "rb").read()
        assert outs[("marked", chunk)] == outs[("bare", chunk)] == outs[("raw", chunk)]
        out = str(tmp_path / ("cut_%s.mp4" % chunk))
        api.transcode(cut, out, args)
        if arbiter.available():
            assert len(arbiter.decode_file(out)) == 30
    assert outs[("bare", None)] == outs[("bare", 6 * 2 * clip.shape[1])]
    if arbiter.available():
        dec = arbiter.decode_file(str(tmp_path / "bare_None.mp4"))
        assert len(dec) == n and arbiter.psnr(dec[n - 1][0], synth.split_planes(clip[n - 1], w, h)[0]) > 32


@pytest.mark.timeout(300)
def test_transcode_many_chunks_cancel_and_timeout(built, tmp_path, monkeypatch):
    """the double-buffered reader: a clip forced into many small chunks gives the same file as one chunk; a cancel flag
    set beforehand and an expired deadline stop the task with the reference's error classes and leave no output"""
    import ctypes as C
    w, h, n = 320, 192, 40
    clip = synth.make_clip(w, h, n, seed=21)
    src = str(tmp_path / "in.y4m")
    with open(src, "wb") as f:
        f.write(b"YUV4MPEG2 W%d H%d F30:1 Ip A1:1 C420\n" % (w, h))
        for fr in clip:
            f.write(b"FRAME\n" + fr.tobytes())
    one = str(tmp_path / "one.mp4")
    api.transcode(src, one, H264_CPU + " -g 5")
    monkeypatch.setenv("VCPENC_CHUNK_BYTES", str(5 * w * h * 3 // 2))
    many = str(tmp_path / "many.mp4")
    api.transcode(src, many, H264_CPU + " -g 5")
    assert open(one, "rb").read() == open(many, "rb").read()
    flag = C.c_int(1)
    out = str(tmp_path / "cancelled.mp4")
    with pytest.raises(api.VcpencError) as e:
        api.transcode(src, out, H264_CPU, cancel=flag)
    assert e.value.code == 6 and not os.path.exists(out)
