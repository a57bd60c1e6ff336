// Input front end for container files: the producer forwards .mp4/.mkv/.avi/.mov/.webm
// (/root/reference/cmd/producer.go:485-488) and the reference's ffmpeg child demuxes and decodes
// them before it encodes (/root/reference/cmd/consumer.go:376-382).  B200 exposes no NVDEC
// headers/libraries in this image, so demux + decode stay on the CPU, exactly as they do inside the
// reference's child process — through the SAME libraries: libavformat / libavcodec, loaded at run
// time with dlopen().  No FFmpeg headers exist here, so only allocation / accessor APIs and the
// long-stable heads of AVFrame, AVPacket, AVFormatContext, AVStream and AVCodecParameters are
// touched (offsets below; checked against the demuxed values at open time).
//
// Library lookup: $VCPENC_FFMPEG_LIBDIR (every *.so* inside is loaded, which is what the hashed
// sonames of the opencv-bundled build need), else the system's libavformat.so.{62..58}.
// This is decode of the INPUT; the encode itself never touches the CPU.
#include <dirent.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "frontend.h"
#include "vcp_algo.h"

namespace vcp {

namespace {

struct AVRational { int num, den; };
struct AVFrameHead { uint8_t* data[8]; int linesize[8]; uint8_t** extended_data; int width, height, nb_samples, format; };
struct AVPacketHead { void* buf; int64_t pts, dts; uint8_t* data; int size; int stream_index; };

struct Lav {
    bool ok = false;
    std::string why;
    // avformat
    int (*avformat_open_input)(void**, const char*, void*, void*) = nullptr;
    int (*avformat_find_stream_info)(void*, void*) = nullptr;
    void (*avformat_close_input)(void**) = nullptr;
    int (*av_find_best_stream)(void*, int, int, int, void*, int) = nullptr;
    int (*av_read_frame)(void*, void*) = nullptr;
    AVRational (*av_guess_frame_rate)(void*, void*, void*) = nullptr;
    // avcodec
    void* (*avcodec_find_decoder)(int) = nullptr;
    void* (*avcodec_alloc_context3)(void*) = nullptr;
    int (*avcodec_parameters_to_context)(void*, void*) = nullptr;
    int (*avcodec_open2)(void*, void*, void*) = nullptr;
    void (*avcodec_free_context)(void**) = nullptr;
    void* (*av_packet_alloc)() = nullptr;
    void (*av_packet_unref)(void*) = nullptr;
    void (*av_packet_free)(void**) = nullptr;
    int (*avcodec_send_packet)(void*, void*) = nullptr;
    int (*avcodec_receive_frame)(void*, void*) = nullptr;
    // avutil
    void* (*av_frame_alloc)() = nullptr;
    void (*av_frame_unref)(void*) = nullptr;
    void (*av_frame_free)(void**) = nullptr;
    int (*av_opt_set_int)(void*, const char*, int64_t, int) = nullptr;
    void (*av_log_set_level)(int) = nullptr;
};

void* find_sym(const std::vector<void*>& hs, const char* name) {
    for (void* h : hs)
        if (void* p = dlsym(h, name)) return p;
    return nullptr;
}

Lav* load_lav() {
    static Lav lav;
    static std::once_flag once;
    std::call_once(once, [] {
        std::vector<void*> hs;
        if (const char* dir = getenv("VCPENC_FFMPEG_LIBDIR")) {
            std::vector<std::string> pending;
            if (DIR* d = opendir(dir)) {
                while (dirent* e = readdir(d))
                    if (strstr(e->d_name, ".so")) pending.push_back(std::string(dir) + "/" + e->d_name);
                closedir(d);
            }
            std::sort(pending.begin(), pending.end());
            for (int round = 0; round < 6 && !pending.empty(); round++) {   // dependencies resolve in some order
                std::vector<std::string> next;
                for (const auto& p : pending) {
                    if (void* h = dlopen(p.c_str(), RTLD_NOW | RTLD_GLOBAL)) hs.push_back(h); else next.push_back(p);
                }
                if (next.size() == pending.size()) break;
                pending.swap(next);
            }
        } else {
            for (const char* stem : {"libavutil.so", "libavcodec.so", "libavformat.so"}) {
                void* h = nullptr;
                for (int v = 64; v >= 55 && !h; v--) h = dlopen((std::string(stem) + "." + std::to_string(v)).c_str(), RTLD_NOW | RTLD_GLOBAL);
                if (!h) h = dlopen(stem, RTLD_NOW | RTLD_GLOBAL);
                if (h) hs.push_back(h);
            }
        }
        if (hs.empty()) { lav.why = "libavformat/libavcodec not found (set VCPENC_FFMPEG_LIBDIR)"; return; }
#define SYM(n) do { *(void**)(&lav.n) = find_sym(hs, #n); if (!lav.n) { lav.why = std::string("missing symbol ") + #n; return; } } while (0)
        SYM(avformat_open_input); SYM(avformat_find_stream_info); SYM(avformat_close_input); SYM(av_find_best_stream);
        SYM(av_read_frame); SYM(av_guess_frame_rate); SYM(avcodec_find_decoder); SYM(avcodec_alloc_context3);
        SYM(avcodec_parameters_to_context); SYM(avcodec_open2); SYM(avcodec_free_context); SYM(av_packet_alloc);
        SYM(av_packet_unref); SYM(av_packet_free); SYM(avcodec_send_packet); SYM(avcodec_receive_frame);
        SYM(av_frame_alloc); SYM(av_frame_unref); SYM(av_frame_free); SYM(av_opt_set_int); SYM(av_log_set_level);
#undef SYM
        lav.av_log_set_level(24);   // AV_LOG_WARNING, the reference runs `-loglevel warning`
        lav.ok = true;
    });
    return &lav;
}

constexpr int AVMEDIA_TYPE_VIDEO = 0, AVMEDIA_TYPE_AUDIO = 1;
constexpr int AVERROR_EOF_ = -541478725;   // FFERRTAG('E','O','F',' ')
// AVPixelFormat values that have been stable since FFmpeg 0.x
constexpr int PIX_YUV420P = 0, PIX_RGB24 = 2, PIX_BGR24 = 3, PIX_YUV422P = 4, PIX_YUV444P = 5, PIX_YUVJ420P = 12,
              PIX_YUVJ422P = 13, PIX_YUVJ444P = 14, PIX_NV12 = 23;

int map_pix_fmt(int f) {
    switch (f) {
    case PIX_YUV420P: case PIX_YUVJ420P: return VCPENC_FMT_YUV420P;
    case PIX_NV12: return VCPENC_FMT_NV12;
    case PIX_RGB24: return VCPENC_FMT_RGB24;
    case PIX_BGR24: return VCPENC_FMT_BGR24;
    case PIX_YUV422P: case PIX_YUVJ422P: return VCPENC_FMT_YUV422P;
    case PIX_YUV444P: case PIX_YUVJ444P: return VCPENC_FMT_YUV444P;
    default: return -1;
    }
}

// copy one decoded picture into the tight layout K1 takes
void pack_frame(const AVFrameHead& f, int fmt, uint8_t* dst) {
    const int w = f.width, h = f.height, cw = (w + 1) / 2, ch = (h + 1) / 2;
    auto plane = [&](int i, int pw, int ph) {
        for (int y = 0; y < ph; y++) { memcpy(dst, f.data[i] + (ptrdiff_t)y * f.linesize[i], (size_t)pw); dst += pw; }
    };
    switch (fmt) {
    case VCPENC_FMT_YUV420P: plane(0, w, h); plane(1, cw, ch); plane(2, cw, ch); break;
    case VCPENC_FMT_NV12: plane(0, w, h); plane(1, 2 * cw, ch); break;
    case VCPENC_FMT_RGB24: case VCPENC_FMT_BGR24: plane(0, 3 * w, h); break;
    case VCPENC_FMT_YUV422P: plane(0, w, h); plane(1, cw, h); plane(2, cw, h); break;
    case VCPENC_FMT_YUV444P: plane(0, w, h); plane(1, w, h); plane(2, w, h); break;
    }
}

struct LavSource : FrameSource {
    Lav* L = nullptr;
    void* fmtctx = nullptr;
    void* dec = nullptr;
    void* pkt = nullptr;
    void* frame = nullptr;
    int vstream = -1;
    bool draining = false, done = false;

    ~LavSource() override {
        if (!L) return;
        if (frame) L->av_frame_free(&frame);
        if (pkt) L->av_packet_free(&pkt);
        if (dec) L->avcodec_free_context(&dec);
        if (fmtctx) L->avformat_close_input(&fmtctx);
    }

    int open(const char* path, bool drop_audio, char* err, size_t errlen) {
        L = load_lav();
        if (!L->ok) { set_err(err, errlen, "container input needs FFmpeg's demuxer/decoder: %s", L->why.c_str()); return VCPENC_E_FORMAT; }
        if (L->avformat_open_input(&fmtctx, path, nullptr, nullptr) < 0) { fmtctx = nullptr; set_err(err, errlen, "cannot open/parse %s", path); return VCPENC_E_FORMAT; }
        if (L->avformat_find_stream_info(fmtctx, nullptr) < 0) { set_err(err, errlen, "no stream information in %s", path); return VCPENC_E_FORMAT; }
        vstream = L->av_find_best_stream(fmtctx, AVMEDIA_TYPE_VIDEO, -1, -1, nullptr, 0);
        if (vstream < 0) { set_err(err, errlen, "no video stream in %s", path); return VCPENC_E_FORMAT; }
        if (!drop_audio && L->av_find_best_stream(fmtctx, AVMEDIA_TYPE_AUDIO, -1, -1, nullptr, 0) >= 0) {
            set_err(err, errlen, "input has an audio stream; audio encoding is not part of this executor (pass -an to drop it)");
            return VCPENC_E_AUDIO;
        }
        // AVFormatContext { av_class, iformat, oformat, priv_data, pb; int ctx_flags; unsigned nb_streams; AVStream** streams }
        const uint8_t* fc = static_cast<const uint8_t*>(fmtctx);
        const unsigned nb_streams = *reinterpret_cast<const unsigned*>(fc + 44);
        if ((unsigned)vstream >= nb_streams || nb_streams > 4096) { set_err(err, errlen, "unexpected AVFormatContext layout (libavformat ABI)"); return VCPENC_E_INTERNAL; }
        void** streams = *reinterpret_cast<void** const*>(fc + 48);
        void* st = streams[vstream];
        // AVStream { av_class; int index, id; AVCodecParameters* codecpar }
        if (*reinterpret_cast<const int*>(static_cast<const uint8_t*>(st) + 8) != vstream) { set_err(err, errlen, "unexpected AVStream layout (libavformat ABI)"); return VCPENC_E_INTERNAL; }
        void* par = *reinterpret_cast<void* const*>(static_cast<const uint8_t*>(st) + 16);
        // AVCodecParameters { int codec_type; int codec_id; ... }
        const int* ip = static_cast<const int*>(par);
        if (ip[0] != AVMEDIA_TYPE_VIDEO) { set_err(err, errlen, "unexpected AVCodecParameters layout (libavcodec ABI)"); return VCPENC_E_INTERNAL; }
        void* codec = L->avcodec_find_decoder(ip[1]);
        if (!codec) { set_err(err, errlen, "no decoder for codec id %d in the loaded libavcodec", ip[1]); return VCPENC_E_FORMAT; }
        dec = L->avcodec_alloc_context3(codec);
        if (!dec || L->avcodec_parameters_to_context(dec, par) < 0) { set_err(err, errlen, "decoder set-up failed"); return VCPENC_E_FORMAT; }
        int threads = (int)std::min<long>(16, std::max<long>(1, sysconf_cores()));
        L->av_opt_set_int(dec, "threads", threads, 0);
        if (L->avcodec_open2(dec, codec, nullptr) < 0) { set_err(err, errlen, "cannot open the decoder"); return VCPENC_E_FORMAT; }
        const AVRational fr = L->av_guess_frame_rate(fmtctx, st, nullptr);
        fps_num = fr.num > 0 ? fr.num : 25; fps_den = fr.den > 0 ? fr.den : 1;
        pkt = L->av_packet_alloc();
        frame = L->av_frame_alloc();
        // decode the first picture to learn size and pixel format
        int rc = next_frame(err, errlen);
        if (rc < 0) return -rc;
        if (rc == 0) { set_err(err, errlen, "input has no decodable video frames"); return VCPENC_E_FORMAT; }
        const AVFrameHead* f = static_cast<const AVFrameHead*>(frame);
        width = f->width; height = f->height;
        fmt = map_pix_fmt(f->format);
        if (fmt < 0) { set_err(err, errlen, "decoded pixel format %d not supported (8-bit 4:2:0 / 4:2:2 / 4:4:4 / nv12 / rgb24)", f->format); return VCPENC_E_FORMAT; }
        have_first = true;
        return VCPENC_OK;
    }

    static long sysconf_cores() { long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? n : 1; }

    bool have_first = false;

    // 1 = a picture is in `frame`, 0 = end of stream, <0 = -error class
    int next_frame(char* err, size_t errlen) {
        for (;;) {
            const int r = L->avcodec_receive_frame(dec, frame);
            if (r >= 0) return 1;
            if (r == AVERROR_EOF_ || done) { done = true; return 0; }
            // needs more input
            if (draining) { done = true; return 0; }
            bool fed = false;
            while (!fed) {
                if (L->av_read_frame(fmtctx, pkt) < 0) { L->avcodec_send_packet(dec, nullptr); draining = true; fed = true; break; }
                const AVPacketHead* ph = static_cast<const AVPacketHead*>(pkt);
                if (ph->stream_index == vstream) {
                    const int s = L->avcodec_send_packet(dec, pkt);
                    fed = true;
                    if (s < 0 && s != -11) { L->av_packet_unref(pkt); set_err(err, errlen, "decoder rejected a packet (%d)", s); return -VCPENC_E_FORMAT; }
                }
                L->av_packet_unref(pkt);
            }
        }
    }

    int read(uint8_t* dst, int max, char* err, size_t errlen) override {
        const size_t fb = fbytes();
        int n = 0;
        while (n < max) {
            if (!have_first) {
                const int r = next_frame(err, errlen);
                if (r < 0) return r;
                if (r == 0) break;
            }
            have_first = false;
            const AVFrameHead* f = static_cast<const AVFrameHead*>(frame);
            if (f->width != width || f->height != height || map_pix_fmt(f->format) != fmt) {
                set_err(err, errlen, "picture size / format changes mid-stream are not supported");
                return -VCPENC_E_FORMAT;
            }
            pack_frame(*f, fmt, dst + (size_t)n * fb);
            L->av_frame_unref(frame);
            n++;
        }
        return n;
    }
};

}  // namespace

int open_container_source(const char* path, bool drop_audio, std::unique_ptr<FrameSource>* out, char* err, size_t errlen) {
    auto s = std::make_unique<LavSource>();
    const int rc = s->open(path, drop_audio, err, errlen);
    if (rc) return rc;
    *out = std::move(s);
    return VCPENC_OK;
}

}  // namespace vcp

// Host-only probe / decode of an input the way vcpenc_transcode will see it: fills the geometry, and
// if `frames` is given decodes up to max_frames pictures into it (tight, *fmt layout).  Used by the
// CPU tests of the front end; needs no GPU.
extern "C" int vcpenc_probe_input(const char* path, int* width, int* height, int* fps_num, int* fps_den, int* fmt,
                                  uint8_t* frames, size_t frames_cap, int max_frames, int* nframes, char* err, size_t errlen) {
    if (!path) { vcp::set_err(err, errlen, "bad arguments"); return VCPENC_E_ARGS; }
    std::unique_ptr<vcp::FrameSource> src;
    int rc = vcp::open_container_source(path, true, &src, err, errlen);
    if (rc) return rc;
    if (width) *width = src->width;
    if (height) *height = src->height;
    if (fps_num) *fps_num = src->fps_num;
    if (fps_den) *fps_den = src->fps_den;
    if (fmt) *fmt = src->fmt;
    int n = 0;
    if (frames && max_frames > 0) {
        const size_t fb = src->fbytes();
        const int can = (int)std::min<size_t>((size_t)max_frames, fb ? frames_cap / fb : 0);
        n = src->read(frames, can, err, errlen);
        if (n < 0) return -n;
    }
    if (nframes) *nframes = n;
    return VCPENC_OK;
}
