// Option grammar of the task's `ffmpeg_args` string.
//
// The producer expands a preset name into a string (/root/reference/internal/config/
// config.go:44-52) and the consumer splits it with strings.Fields — whitespace only, no
// quoting (/root/reference/cmd/consumer.go:378).  This parser understands exactly those
// tokens plus a few genuine ffmpeg option names used as knobs (-g, -coder, -slices, -qp,
// -r, -s, -pix_fmt), so the same task string stays valid for a stock ffmpeg.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/vcpenc.h"
#include "host_util.h"

namespace vcp {

void set_err(char* err, size_t errlen, const char* fmt, ...) {
    if (!err || !errlen) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err, errlen, fmt, ap);
    va_end(ap);
}

// "10M", "128k", "2500000" -> bits per second; -1 on error
static long parse_rate(const char* s) {
    char* end = nullptr;
    double v = strtod(s, &end);
    if (end == s || v < 0) return -1;
    if (*end == 'k' || *end == 'K') { v *= 1000; end++; }
    else if (*end == 'm' || *end == 'M') { v *= 1000000; end++; }
    else if (*end == 'g' || *end == 'G') { v *= 1000000000; end++; }
    if (*end != 0) return -1;
    if (v > 2147483647.0) return -1;   // params hold int32 bits per second
    return (long)v;
}

static bool parse_int(const char* s, int* out) {
    char* end = nullptr;
    long v = strtol(s, &end, 10);
    if (end == s || *end) return false;
    *out = (int)v;
    return true;
}

}  // namespace vcp

using namespace vcp;

extern "C" int vcpenc_parse_args(int argc, const char* const* argv, vcpenc_params* p, char* err, size_t errlen) {
    if (!p || (argc > 0 && !argv)) { set_err(err, errlen, "bad arguments"); return VCPENC_E_ARGS; }
    vcpenc_default_params(p);
    p->slices = 0;     // unless -slices is given the encoder chooses (vcp_auto_slices)
    p->entropy = -1;   // unless -coder / -profile:v says otherwise: the codec's default (CABAC for libx264 / nvenc)
    p->transform8x8 = -1;   // likewise: High profile tools unless -profile:v baseline / main
    bool have_codec = false, have_crf = false, have_qp = false;
    int crf = 23;
    int sao = -1, subpel = -1;   // HEVC tools from -x265-params (-1: not given)
    for (int i = 0; i < argc; i++) {
        const std::string t = argv[i];
        auto need = [&](const char** v) -> bool {
            if (i + 1 >= argc) { set_err(err, errlen, "option %s needs a value", t.c_str()); return false; }
            *v = argv[++i];
            return true;
        };
        const char* v = nullptr;
        if (t == "-c:v" || t == "-vcodec" || t == "-codec:v") {
            if (!need(&v)) return VCPENC_E_ARGS;
            const std::string c = v;
            if (c == "libx264" || c == "h264_nvenc" || c == "h264" || c == "libopenh264") p->codec = VCPENC_CODEC_H264;
            else if (c == "libx265" || c == "hevc_nvenc" || c == "hevc") p->codec = VCPENC_CODEC_HEVC;
            else if (c == "copy") { set_err(err, errlen, "stream copy is not an encode"); return VCPENC_E_NOTENCODE; }
            else { set_err(err, errlen, "unknown video codec '%s'", v); return VCPENC_E_ARGS; }
            have_codec = true;
        } else if (t == "-c" || t == "-codec") {
            if (!need(&v)) return VCPENC_E_ARGS;
            if (!strcmp(v, "copy")) { set_err(err, errlen, "stream copy is not an encode"); return VCPENC_E_NOTENCODE; }
            set_err(err, errlen, "unsupported -c %s", v);
            return VCPENC_E_ARGS;
        } else if (t == "-vn") {
            set_err(err, errlen, "-vn: no video to encode");
            return VCPENC_E_NOTENCODE;
        } else if (t == "-preset") {
            if (!need(&v)) return VCPENC_E_ARGS;
            const std::string s = v;
            if (s == "ultrafast" || s == "superfast" || s == "veryfast" || s == "faster" || s == "fast" || s == "p1" || s == "p2" || s == "p3") p->effort = 0;
            else if (s == "medium" || s == "p4" || s == "p5") p->effort = 1;
            else if (s == "slow" || s == "slower" || s == "veryslow" || s == "placebo" || s == "p6" || s == "p7") p->effort = 2;
            else { set_err(err, errlen, "unknown preset '%s'", v); return VCPENC_E_ARGS; }
        } else if (t == "-profile:v" || t == "-profile") {
            if (!need(&v)) return VCPENC_E_ARGS;
            // baseline has no CABAC; main / high default to it (an explicit -coder still wins)
            if (p->entropy < 0 && !strcmp(v, "baseline")) p->entropy = 0;
            p->transform8x8 = !strncmp(v, "high", 4) ? 1 : 0;   // baseline, main: 4x4 transform only
        } else if (t == "-tune" || t == "-level" || t == "-threads" || t == "-refs" ||
                   t == "-c:a" || t == "-acodec" || t == "-ar" || t == "-ac" || t == "-f" ||
                   t == "-rc" || t == "-rc-lookahead" || t == "-x264-params" || t == "-x264opts") {
            if (!need(&v)) return VCPENC_E_ARGS;  // accepted, no effect on this encoder
        } else if (t == "-b:a" || t == "-ab") {
            if (!need(&v)) return VCPENC_E_ARGS;
            const long r = parse_rate(v);
            if (r <= 0 || r > 2000000) { set_err(err, errlen, "bad audio bitrate '%s'", v); return VCPENC_E_ARGS; }
            p->audio_bitrate = (int32_t)r;
        } else if (t == "-x265-params") {
            // x265's own option string (key=value pairs joined by ':'): sao / no-sao and the sub-sample search are understood
            if (!need(&v)) return VCPENC_E_ARGS;
            const std::string xs = std::string(":") + v + ":";
            if (xs.find(":sao=1:") != std::string::npos || xs.find(":sao:") != std::string::npos) sao = 1;
            if (xs.find(":sao=0:") != std::string::npos || xs.find(":no-sao:") != std::string::npos || xs.find(":no-sao=1:") != std::string::npos) sao = 0;
            {   // subme=N: 0 full samples only, 1 half samples, 2..4 quarter samples (candidates ranked by the half-sample proxy), 5.. ranked exactly
                const size_t at = xs.find(":subme=");
                if (at != std::string::npos) { const int lv = atoi(xs.c_str() + at + 7); subpel = lv <= 0 ? 0 : lv == 1 ? 1 : lv < 5 ? 2 : 3; }
            }
        } else if (t == "-an") {
            p->drop_audio = 1;
        } else if (t == "-sn" || t == "-dn" || t == "-y" || t == "-hide_banner" || t == "-nostdin") {
            // accepted
        } else if (t == "-loglevel" || t == "-v") {
            if (!need(&v)) return VCPENC_E_ARGS;
        } else if (t == "-crf") {
            if (!need(&v)) return VCPENC_E_ARGS;
            double d = atof(v);
            if (d < 0 || d > 51) { set_err(err, errlen, "bad -crf %s", v); return VCPENC_E_ARGS; }
            crf = (int)(d + 0.5); have_crf = true;
        } else if (t == "-qp" || t == "-cqp") {
            int q;
            if (!need(&v) || !parse_int(v, &q) || q < 0 || q > 51) { set_err(err, errlen, "bad %s", t.c_str()); return VCPENC_E_ARGS; }
            p->qp_p = q; p->qp_i = q > 2 ? q - 2 : 0; have_qp = true;
        } else if (t == "-b:v" || t == "-b") {
            if (!need(&v)) return VCPENC_E_ARGS;
            long r = parse_rate(v);
            if (r <= 0) { set_err(err, errlen, "bad bitrate '%s'", v); return VCPENC_E_ARGS; }
            p->bitrate = (int32_t)r; p->rc_mode = VCPENC_RC_ABR;
        } else if (t == "-maxrate") {
            if (!need(&v)) return VCPENC_E_ARGS;
            long r = parse_rate(v);
            if (r <= 0) { set_err(err, errlen, "bad -maxrate '%s'", v); return VCPENC_E_ARGS; }
            p->maxrate = (int32_t)r;
        } else if (t == "-bufsize") {
            if (!need(&v)) return VCPENC_E_ARGS;
            long r = parse_rate(v);
            if (r <= 0) { set_err(err, errlen, "bad -bufsize '%s'", v); return VCPENC_E_ARGS; }
            p->bufsize = (int32_t)r;
        } else if (t == "-movflags") {
            if (!need(&v)) return VCPENC_E_ARGS;
            if (strstr(v, "faststart")) p->faststart = 1;
        } else if (t == "-g") {
            int gop;
            if (!need(&v) || !parse_int(v, &gop) || gop < 1) { set_err(err, errlen, "bad -g"); return VCPENC_E_ARGS; }
            p->gop = gop;
        } else if (t == "-bf") {
            int bf;
            if (!need(&v) || !parse_int(v, &bf) || bf < 0) { set_err(err, errlen, "bad -bf"); return VCPENC_E_ARGS; }
            // B-frames are not produced; 0 is the only exact match, others are accepted as a hint
        } else if (t == "-coder") {
            if (!need(&v)) return VCPENC_E_ARGS;
            const std::string s = v;
            if (s == "0" || s == "vlc" || s == "cavlc") p->entropy = 0;
            else if (s == "1" || s == "ac" || s == "cabac") p->entropy = 1;
            else { set_err(err, errlen, "bad -coder '%s'", v); return VCPENC_E_ARGS; }
        } else if (t == "-slices") {
            int n;
            if (!need(&v) || !parse_int(v, &n) || n < 1) { set_err(err, errlen, "bad -slices"); return VCPENC_E_ARGS; }
            p->slices = n;
        } else if (t == "-r" || t == "-framerate") {
            if (!need(&v)) return VCPENC_E_ARGS;
            int a = 0, b = 1;
            if (sscanf(v, "%d/%d", &a, &b) < 1 || a <= 0 || b <= 0) { set_err(err, errlen, "bad frame rate '%s'", v); return VCPENC_E_ARGS; }
            p->fps_num = a; p->fps_den = b;
        } else if (t == "-s" || t == "-video_size") {
            if (!need(&v)) return VCPENC_E_ARGS;
            int w = 0, h = 0;
            if (sscanf(v, "%dx%d", &w, &h) != 2 || w <= 0 || h <= 0) { set_err(err, errlen, "bad size '%s'", v); return VCPENC_E_ARGS; }
            p->in_width = w; p->in_height = h;
        } else if (t == "-vf" || t == "-filter:v") {
            // only the filter the presets could need: scale=W:H (-1 / -2 keep the aspect ratio;
            // resolved against the input size by vcpenc_transcode)
            if (!need(&v)) return VCPENC_E_ARGS;
            int w = 0, h = 0;
            if (sscanf(v, "scale=%d:%d", &w, &h) != 2 || w == 0 || h == 0 || w < -2 || h < -2 || (w < 0 && h < 0)) {
                set_err(err, errlen, "unsupported filter '%s' (scale=W:H only)", v);
                return VCPENC_E_ARGS;
            }
            p->width = w; p->height = h;
        } else if (t == "-pix_fmt") {
            if (!need(&v)) return VCPENC_E_ARGS;
            if (strcmp(v, "yuv420p")) { set_err(err, errlen, "output pix_fmt '%s' unsupported (yuv420p only)", v); return VCPENC_E_ARGS; }
        } else {
            set_err(err, errlen, "unrecognised option '%s'", t.c_str());
            return VCPENC_E_ARGS;
        }
    }
    (void)have_codec;
    if (p->entropy < 0) p->entropy = 1;   // x264 and NVENC both default to CABAC
    if (p->transform8x8 < 0) p->transform8x8 = 1;   // ... and to High profile
    if (p->codec == VCPENC_CODEC_HEVC) {
        // libx265 / hevc_nvenc search sub-sample positions: quarter samples here too, half samples in the fast -preset
        // tiers (like the H.264 refine).  SAO is implemented and bit-exact but its first kernel costs ~45 us per 1080p
        // picture (profiles/r01_notes.md), so it is opt-in: -x265-params sao=1
        p->hevc_subpel = subpel >= 0 ? subpel : (p->effort == 0 ? 1 : 2);
        p->hevc_sao = sao > 0 ? 1 : 0;
    }
    if (have_crf && !have_qp) {
        // constant quality: one QP per picture type (x264's default ipratio 1.4 ~ 3 QP)
        p->qp_p = crf + 1 > 51 ? 51 : crf + 1;
        p->qp_i = p->qp_p >= 3 ? p->qp_p - 3 : 0;
        if (p->rc_mode != VCPENC_RC_ABR) p->rc_mode = VCPENC_RC_CQP;
    }
    return VCPENC_OK;
}
