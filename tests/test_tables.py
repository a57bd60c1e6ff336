"""Normative H.264 constants typed in csrc/h264_tables.h must equal the ones inside the
reference FFmpeg decoder (libavcodec .rodata) — a mechanical cross-check of hand-typed tables."""
import glob
import os
import re

import pytest

from video_codec_pipeline_b200 import arbiter

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "video_codec_pipeline_b200", "csrc", "h264_tables.h")


def parse_tables():
    src = open(HDR).read()
    out = {}
    for m in re.finditer(r"VCP_TAB\s+\w+\s+(\w+)((?:\[\d+\])+)\s*=\s*\{(.*?)\};", src, re.S):
        out[m.group(1)] = [int(x) for x in re.findall(r"-?\d+", m.group(3))]
    return out


def libavcodec_bytes():
    d = arbiter._find_libdir()
    if d is None:
        pytest.skip("bundled libavcodec not present")
    path = sorted(glob.glob(os.path.join(d, "libavcodec-*.so*")))[0]
    return open(path, "rb").read()


def test_table_shapes():
    t = parse_tables()
    assert len(t["vcp_zigzag4x4"]) == 16 and sorted(t["vcp_zigzag4x4"]) == list(range(16))
    assert len(t["vcp_coeff_token_len"]) == 4 * 68 and len(t["vcp_coeff_token_bits"]) == 4 * 68
    assert len(t["vcp_total_zeros_len"]) == 15 * 16
    assert len(t["vcp_alpha_tab"]) == 52 and len(t["vcp_beta_tab"]) == 52 and len(t["vcp_tc0_tab"]) == 156
    assert sorted(t["vcp_cbp_to_golomb_intra"]) == list(range(48))
    assert sorted(t["vcp_cbp_to_golomb_inter"]) == list(range(48))


def test_vlc_tables_are_prefix_free():
    t = parse_tables()

    def codes(lens, bits):
        return [(l, b) for l, b in zip(lens, bits) if l]

    def prefix_free(cs):
        s = sorted(format(b, "0%db" % l) for l, b in cs)
        return all(not s[i + 1].startswith(s[i]) for i in range(len(s) - 1))

    for tab in range(4):
        assert prefix_free(codes(t["vcp_coeff_token_len"][68 * tab:68 * tab + 68], t["vcp_coeff_token_bits"][68 * tab:68 * tab + 68]))
    assert prefix_free(codes(t["vcp_chroma_dc_coeff_token_len"], t["vcp_chroma_dc_coeff_token_bits"]))
    for tc in range(15):
        assert prefix_free(codes(t["vcp_total_zeros_len"][16 * tc:16 * tc + 16], t["vcp_total_zeros_bits"][16 * tc:16 * tc + 16]))
    for z in range(7):
        assert prefix_free(codes(t["vcp_run_len"][16 * z:16 * z + 16], t["vcp_run_bits"][16 * z:16 * z + 16]))


def test_tables_match_ffmpeg_decoder_rodata():
    t = parse_tables()
    blob = libavcodec_bytes()

    def present(seq):
        return blob.find(bytes(seq)) >= 0

    assert present(t["vcp_coeff_token_len"])
    assert present(t["vcp_coeff_token_bits"])
    assert present(t["vcp_chroma_dc_coeff_token_len"])
    assert present(t["vcp_chroma_dc_coeff_token_bits"])
    assert present(t["vcp_total_zeros_len"])      # ffmpeg stores 16 rows; ours are the first 15
    assert present(t["vcp_total_zeros_bits"])
    assert present(t["vcp_chroma_dc_total_zeros_len"]) and present(t["vcp_chroma_dc_total_zeros_bits"])
    assert present(t["vcp_run_len"][:111]) and present(t["vcp_run_bits"][:111])
    assert present(t["vcp_alpha_tab"][16:]) and present(t["vcp_beta_tab"][16:])
    tc0 = t["vcp_tc0_tab"]
    assert present(b"".join(bytes([255] + tc0[3 * i:3 * i + 3]) for i in range(52)))
    assert present(t["vcp_chroma_qp"][30:])
    assert present(t["vcp_zigzag4x4"]) or present([0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15])
    # Table 9-4: ffmpeg stores codeNum -> cbp; ours is the inverse
    for name, inv in (("intra", t["vcp_cbp_to_golomb_intra"]), ("inter", t["vcp_cbp_to_golomb_inter"])):
        fwd = [0] * 48
        for cbp, code in enumerate(inv):
            fwd[code] = cbp
        assert present(fwd), name
    deq = t["vcp_dequant_v"]   # [6][3] = (a,b,c); ffmpeg: {a, c, b} per row
    assert present(sum(([deq[3 * i], deq[3 * i + 2], deq[3 * i + 1]] for i in range(6)), []))
