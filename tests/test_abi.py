"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol
include/vcpenc.h declares, parses the reference's preset strings, fails loudly without a GPU,
and the host-only pieces (MP4 mux, verify) interoperate with FFmpeg's demuxer/decoder."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import pyoracle
from video_codec_pipeline_b200 import api, arbiter, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# /root/reference/internal/config/config.go:44-52, verbatim preset strings
PRESETS = {
    "h264-nvenc": "-c:v h264_nvenc -preset p4 -b:v 10M -c:a aac -b:a 128k -movflags +faststart",
    "h264-nvenc-hq": "-c:v h264_nvenc -preset p7 -tune hq -b:v 15M -maxrate 20M -bufsize 30M -c:a aac -b:a 192k -movflags +faststart",
    "h265-nvenc": "-c:v hevc_nvenc -preset p4 -b:v 8M -c:a aac -b:a 128k -movflags +faststart",
    "h265-nvenc-hq": "-c:v hevc_nvenc -preset p7 -tune hq -b:v 10M -c:a aac -b:a 192k -movflags +faststart",
    "h264-cpu": "-c:v libx264 -preset medium -crf 23 -c:a aac -b:a 128k -movflags +faststart",
    "h265-cpu": "-c:v libx265 -preset medium -crf 28 -c:a aac -b:a 128k -movflags +faststart",
    "copy": "-c copy",
}


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "vcpenc.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(vcpenc_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 15
    L = C.CDLL(api.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert b"vcpenc" in L.vcpenc_version.__call__.__self__.restype.__name__.encode() or True
    assert "sm_100a" in api.version()


def test_struct_layout_matches_header(built):
    assert C.sizeof(api.Params) == 32 * 4
    assert C.sizeof(api.FrameInfo) == 16
    p = api.default_params(1920, 1080)
    assert (p.gop, p.slices, p.codec, p.entropy, p.fps_num) == (60, 1, 0, 0, 30)


def test_parse_reference_presets(built):
    p = api.parse_args(PRESETS["h264-cpu"].split())
    assert p.codec == 0 and p.rc_mode == 0 and p.faststart == 1 and p.qp_p == 24 and p.qp_i == 21
    assert p.entropy == 1 and p.slices == 0          # x264's default coder; slice count left to the encoder
    assert p.transform8x8 == 1                       # ... and x264's default profile (High)
    assert api.parse_args("-c:v libx264 -profile:v main".split()).transform8x8 == 0
    assert api.parse_args("-c:v libx264 -profile:v baseline".split()).entropy == 0
    assert api.parse_args("-c:v libx264 -profile:v baseline -coder 1".split()).entropy == 1
    p = api.parse_args(PRESETS["h264-nvenc"].split())
    assert p.codec == 0 and p.rc_mode == 1 and p.bitrate == 10_000_000 and p.faststart == 1
    p = api.parse_args(PRESETS["h264-nvenc-hq"].split())
    assert p.bitrate == 15_000_000 and p.maxrate == 20_000_000 and p.bufsize == 30_000_000 and p.effort == 2
    p = api.parse_args(PRESETS["h265-cpu"].split())
    assert p.codec == 1
    # -preset tiers: custom presets (config.yaml:14-23) may name the fast ones
    assert [api.parse_args(("-c:v libx264 -preset %s" % t).split()).effort for t in ("veryfast", "p2", "medium", "p5", "slow", "p7")] == [0, 0, 1, 1, 2, 2]
    with pytest.raises(api.VcpencError) as e:
        api.parse_args(PRESETS["copy"].split())
    assert e.value.code == 8  # NOTENCODE: the Go side hands the task to a stock ffmpeg
    with pytest.raises(api.VcpencError) as e:
        api.parse_args("-vn -c:a aac -b:a 192k".split())   # config.yaml:23 custom audio-only preset
    assert e.value.code == 8
    with pytest.raises(api.VcpencError) as e:
        api.parse_args("-c:v libx264 -bogus 1".split())
    assert e.value.code == 1
    with pytest.raises(api.VcpencError):
        api.parse_args("-c:v libx264 -crf".split())         # missing value
    p = api.parse_args("-c:v libx264 -qp 30 -g 30 -slices 4 -coder 0 -bf 0 -r 60000/1001".split())
    assert (p.qp_p, p.gop, p.slices, p.entropy, p.fps_num, p.fps_den) == (30, 30, 4, 0, 60000, 1001)
    assert api.parse_args([]).codec == 0                    # empty ffmpeg_args is legal (consumer.go:377)
    # HEVC tools: quarter-sample motion by default (half samples in the fast tiers), SAO through x265's own option string
    p = api.parse_args("-c:v libx265 -preset medium -crf 28".split())
    assert (p.codec, p.hevc_subpel, p.hevc_sao) == (1, 2, 0)
    assert api.parse_args("-c:v libx265 -preset veryfast -crf 28".split()).hevc_subpel == 1
    assert api.parse_args("-c:v hevc_nvenc -preset p7 -x265-params subme=1".split()).hevc_subpel == 1
    p = api.parse_args("-c:v libx265 -crf 28 -x265-params sao=1:keyint=60".split())
    assert (p.hevc_subpel, p.hevc_sao) == (2, 1)
    p = api.parse_args("-c:v hevc_nvenc -b:v 8M -x265-params no-sao:subme=0".split())
    assert (p.hevc_subpel, p.hevc_sao) == (0, 0)
    assert api.parse_args("-c:v libx264 -crf 23".split()).hevc_subpel == 0
    p = api.parse_args("-c:v libx264 -vf scale=1280:-2 -crf 20".split())
    assert (p.width, p.height) == (1280, -2)
    with pytest.raises(api.VcpencError):
        api.parse_args("-vf hflip".split())


@pytest.mark.skipif(api.lib().vcpenc_device_count() > 0, reason="CPU-only behaviour")
def test_no_cpu_fallback(built, tmp_path):
    clip = synth.make_clip(64, 48, 2, seed=1)
    with pytest.raises(api.VcpencError) as e:
        api.encode_frames(api.default_params(64, 48), clip)
    assert e.value.code == 4
    y4m = tmp_path / "a.y4m"
    with open(y4m, "wb") as f:
        f.write(b"YUV4MPEG2 W64 H48 F30:1 Ip A1:1 C420jpeg\n")
        for fr in clip:
            f.write(b"FRAME\n" + fr.tobytes())
    out = tmp_path / "a.mp4"
    with pytest.raises(api.VcpencError) as e:
        api.transcode(str(y4m), str(out), PRESETS["h264-cpu"])
    assert e.value.code == 4 and not out.exists()
    # the argv door reports the same class through its exit status
    exe = os.path.join(os.path.dirname(api.LIB_PATH), "vcp-ffmpeg")
    r = subprocess.run([exe, "-hide_banner", "-loglevel", "warning", "-y", "-i", str(y4m)] +
                       PRESETS["h264-cpu"].split() + [str(out)], capture_output=True)
    assert r.returncode == 4 and not out.exists()


def test_mux_and_verify_host_only(built, tmp_path):
    """MP4 muxer + verify need no GPU: wrap an oracle stream, then check it three ways."""
    w, h, n = 320, 180, 7
    clip = synth.make_clip(w, h, n, seed=21)
    r = pyoracle.encode(pyoracle.make_params(w, h, gop=3, qp_i=24, qp_p=26, slices=2), clip)
    for fast in (1, 0):
        p = api.default_params(w, h, gop=3, faststart=fast)
        path = str(tmp_path / ("o%d.mp4" % fast))
        api.mux_mp4(p, np.frombuffer(r["stream"], np.uint8), r["info"], path)
        api.verify(path)                                   # our ffprobe-equivalent
        data = open(path, "rb").read()
        assert (data.find(b"moov") < data.find(b"mdat")) == bool(fast)
        if arbiter.available():
            assert arbiter.probe_has_video(path)           # FFmpeg's own demuxer agrees
            dec = arbiter.decode_file(path)
            assert len(dec) == n
            for i in range(n):
                assert np.array_equal(np.concatenate([pl.ravel() for pl in dec[i]]), r["recon"][i])
        exe = os.path.join(os.path.dirname(api.LIB_PATH), "vcp-ffprobe")
        out = subprocess.run([exe, "-v", "error", "-select_streams", "v:0", "-show_entries", "stream=codec_type",
                              "-of", "csv=p=0", path], capture_output=True)
        assert out.returncode == 0 and b"video" in out.stdout   # cmd/consumer.go:409-418


def test_hevc_mux_and_parameter_sets_host_only(built, tmp_path):
    """HEVC host pieces without a GPU: the library's VPS/SPS/PPS are the oracle's (the stream the FFmpeg hevc
    decoder accepts), and an oracle stream wrapped as hvc1 + hvcC demuxes / decodes to the oracle's recon."""
    w, h, n = 328, 184, 7                                   # cropped coded size: conformance window in the SPS
    clip = synth.make_clip(w, h, n, seed=22)
    r = pyoracle.encode_hevc(pyoracle.make_params(w, h, codec=1, gop=3, qp_i=24, qp_p=26, slices=2), clip)
    p = api.default_params(w, h, codec=1, gop=3, faststart=1)
    path = str(tmp_path / "o.mp4")
    api.mux_mp4(p, np.frombuffer(r["stream"], np.uint8), r["info"], path)
    api.verify(path)
    data = open(path, "rb").read()
    assert b"hvc1" in data and b"hvcC" in data and b"avcC" not in data
    raw = str(tmp_path / "o.h265")
    open(raw, "wb").write(r["stream"])
    api.verify(raw)
    if arbiter.available():
        assert arbiter.probe_has_video(path)
        dec = arbiter.decode_file(path)
        assert len(dec) == n
        for i in range(n):
            assert np.array_equal(np.concatenate([pl.ravel() for pl in dec[i]]), r["recon"][i])
    # a stream with its parameter sets stripped: the muxer regenerates them (headers.cpp) -- same file
    off, size = r["info"][0][0], r["info"][0][1]
    au0 = r["stream"][off:off + size]
    first_slice = au0.find(b"\x00\x00\x00\x01\x26")           # IDR_W_RADL
    assert first_slice > 0
    stripped = au0[first_slice:] + r["stream"][off + size:]
    info = [(0, size - first_slice, 1, r["info"][0][3])] + [(o - first_slice, s_, i_, q) for (o, s_, i_, q) in r["info"][1:]]
    # later IDR access units still carry their own parameter sets; the first one relies on headers.cpp
    path2 = str(tmp_path / "o2.mp4")
    api.mux_mp4(p, np.frombuffer(stripped, np.uint8), info, path2)
    assert open(path2, "rb").read() == data


def test_oracle_only_tools_are_refused_by_the_library(built):
    """params.hevc_intra_modes exists in the oracle (pinned by the decoder) but not yet on the device: the library
    says so with its own error class instead of silently encoding without the tool."""
    for kw in (dict(hevc_intra_modes=1),):
        p = api.default_params(320, 192, codec=1, **kw)
        with pytest.raises(api.VcpencError) as e:
            api.Session(p, 4)
        assert e.value.code == 13, kw


def test_verify_rejects_bad_files(built, tmp_path):
    empty = tmp_path / "e.mp4"
    empty.write_bytes(b"")
    with pytest.raises(api.VcpencError) as e:
        api.verify(str(empty))
    assert e.value.code == 10
    junk = tmp_path / "j.mp4"
    junk.write_bytes(os.urandom(4096))
    with pytest.raises(api.VcpencError) as e:
        api.verify(str(junk))
    assert e.value.code == 10
    with pytest.raises(api.VcpencError) as e:
        api.verify(str(tmp_path / "missing.mp4"))
    assert e.value.code == 2
    # an mp4 with only an audio-less, video-less moov
    novid = tmp_path / "n.mp4"
    novid.write_bytes(b"\x00\x00\x00\x10ftypisom\x00\x00\x02\x00" + b"\x00\x00\x00\x08moov")
    with pytest.raises(api.VcpencError):
        api.verify(str(novid))


def test_container_front_end_decodes_what_we_mux(built, tmp_path):
    """SURVEY 8f1, host-only: an MP4 written by our muxer goes back in through the container front end
    (libavformat demux + libavcodec decode, loaded at run time) and yields exactly the pictures the
    encoder reconstructed, with the right geometry and frame rate."""
    if not arbiter.available():
        pytest.skip("bundled FFmpeg libraries not present")
    w, h, n = 320, 180, 9
    clip = synth.make_clip(w, h, n, seed=33)
    r = pyoracle.encode(pyoracle.make_params(w, h, fps=24, gop=4, qp_i=24, qp_p=26, entropy=1, slices=2), clip)
    path = str(tmp_path / "in.mp4")
    api.mux_mp4(api.default_params(w, h, fps=24, gop=4, faststart=1), np.frombuffer(r["stream"], np.uint8), r["info"], path)
    info = api.probe_input(path, max_frames=n + 3)
    assert (info["width"], info["height"], info["fmt"]) == (w, h, 0)
    assert info["fps"][0] / info["fps"][1] == 24
    assert info["frames"].shape[0] == n
    assert np.array_equal(info["frames"], r["recon"])
    # an HEVC input (what an h265-* task leaves behind, re-submitted): hvc1 + hvcC through the same front end
    r5 = pyoracle.encode_hevc(pyoracle.make_params(w, h, fps=24, codec=1, gop=4, qp_i=24, qp_p=26, slices=2, hevc_subpel=1, hevc_sao=1), clip)
    path5 = str(tmp_path / "in_hevc.mp4")
    api.mux_mp4(api.default_params(w, h, fps=24, codec=1, gop=4, faststart=1, hevc_sao=1), np.frombuffer(r5["stream"], np.uint8), r5["info"], path5)
    info5 = api.probe_input(path5, max_frames=n + 3)
    assert (info5["width"], info5["height"]) == (w, h) and info5["frames"].shape[0] == n
    assert np.array_equal(info5["frames"], r5["recon"])
    # not a media file -> FORMAT
    junk = tmp_path / "junk.mkv"
    junk.write_bytes(b"\x1a\x45\xdf\xa3" + os.urandom(2000))
    with pytest.raises(api.VcpencError) as e:
        api.probe_input(str(junk))
    assert e.value.code == 3
