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
#include <memory>
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
struct AVChannelLayoutPub { int order; int nb_channels; uint64_t mask; void* opaque; };   // public, fixed layout since lavu 57.24
struct AVOptionHead { const char* name; const char* help; int offset; int type; };

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
    // audio (optional: without them an input with audio is handed back as VCPENC_E_AUDIO)
    bool audio_ok = false;
    void* (*avcodec_find_encoder_by_name)(const char*) = nullptr;
    int (*avcodec_send_frame)(void*, const void*) = nullptr;
    int (*avcodec_receive_packet)(void*, void*) = nullptr;
    int (*av_opt_get_int)(void*, const char*, int, int64_t*) = nullptr;
    const void* (*av_opt_find)(void*, const char*, const char*, int, int) = nullptr;
    int (*av_opt_get_chlayout)(void*, const char*, int, void*) = nullptr;
    int (*av_opt_set_chlayout)(void*, const char*, const void*, int) = nullptr;
    void (*av_channel_layout_default)(void*, int) = nullptr;
    int (*av_frame_ref)(void*, const void*) = nullptr;
    int (*swr_alloc_set_opts2)(void**, const void*, int, int, const void*, int, int, int, void*) = nullptr;
    int (*swr_init)(void*) = nullptr;
    int (*swr_convert)(void*, uint8_t**, int, const uint8_t**, int) = nullptr;
    int (*swr_get_out_samples)(void*, int) = nullptr;
    void (*swr_free)(void**) = nullptr;
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
#define OSYM(n) (*(void**)(&lav.n) = find_sym(hs, #n))
        lav.audio_ok = OSYM(avcodec_find_encoder_by_name) && OSYM(avcodec_send_frame) && OSYM(avcodec_receive_packet) && OSYM(av_opt_get_int) &&
                       OSYM(av_opt_find) && OSYM(av_opt_get_chlayout) && OSYM(av_opt_set_chlayout) && OSYM(av_channel_layout_default) &&
                       OSYM(av_frame_ref) && OSYM(swr_alloc_set_opts2) && OSYM(swr_init) && OSYM(swr_convert) && OSYM(swr_get_out_samples) && OSYM(swr_free);
#undef OSYM
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


// ---- audio -----------------------------------------------------------------------------------------------
// Every encode preset carries `-c:a aac -b:a Nk` (/root/reference/internal/config/config.go:45-50), so an input
// with an audio stream must come out with an AAC track.  Two ways, both on the CPU like the reference's child:
//   * the input is AAC-LC with an AudioSpecificConfig: its access units are copied (no generation loss);
//   * anything else is decoded, converted to planar float by libswresample (format, and rate / layout when the
//     AAC encoder does not take the input's) and encoded by libavcodec's native `aac` encoder at -b:a.
// Header-less: AVCodecContext fields are set through AVOptions ("b", "ar", "ch_layout"); sample_fmt has no
// option, its offset is taken from its neighbours' ("ar" sits right in front of it in every libavcodec since 5.1).
// Frames for the encoder are a reference of a decoded frame (which carries a valid layout / rate / refcounted
// buffers) whose data pointers, sample count and format are pointed at our own planar float block.
constexpr int CODEC_ID_AAC = 0x15002;
constexpr int SAMPLE_FMT_FLTP = 8;
constexpr int AAC_FRAME = 1024;

int aac_rate_index(int rate) {
    static const int r[13] = {96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000, 12000, 11025, 8000, 7350};
    for (int i = 0; i < 13; i++) if (r[i] == rate) return i;
    return -1;
}

struct LavAudio {
    Lav* L = nullptr;
    AudioTrack* out = nullptr;
    std::mutex* mu = nullptr;
    int stream = -1;
    bool copy = false;
    void *dec = nullptr, *enc = nullptr, *swr = nullptr, *dframe = nullptr, *carrier = nullptr, *opkt = nullptr;
    bool configured = false;
    int in_fmt = -1, in_rate = 0;
    std::vector<std::vector<float>> fifo;   // planar, one vector per output channel
    std::vector<float> scratch;
    int64_t sent = 0;                       // samples handed to the encoder
    int bitrate = 128000;

    ~LavAudio() {
        if (!L) return;
        if (carrier) L->av_frame_free(&carrier);
        if (dframe) L->av_frame_free(&dframe);
        if (opkt) L->av_packet_free(&opkt);
        if (enc) L->avcodec_free_context(&enc);
        if (dec) L->avcodec_free_context(&dec);
        if (swr) L->swr_free(&swr);
    }

    void emit(const uint8_t* d, int n) {
        std::lock_guard<std::mutex> lk(*mu);
        out->data.insert(out->data.end(), d, d + n);
        out->sizes.push_back((uint32_t)n);
        out->frames_total++;
    }

    // par: the stream's AVCodecParameters { int codec_type; int codec_id; uint32_t codec_tag; uint8_t* extradata; int extradata_size; ... }
    int open(Lav* lav, void* par, int stream_index, int bitrate_, AudioTrack* o, std::mutex* m, char* err, size_t errlen) {
        L = lav; out = o; mu = m; stream = stream_index;
        if (bitrate_ > 0) bitrate = bitrate_;
        const uint8_t* pb = static_cast<const uint8_t*>(par);
        const int codec_id = reinterpret_cast<const int*>(pb)[1];
        const uint8_t* extra = *reinterpret_cast<uint8_t* const*>(pb + 16);
        const int extra_size = *reinterpret_cast<const int*>(pb + 24);
        if (codec_id == CODEC_ID_AAC && extra && extra_size >= 2 && extra_size < 64) {
            // AudioSpecificConfig: 5 bits object type, 4 bits frequency index, 4 bits channel configuration
            const int aot = extra[0] >> 3, fi = ((extra[0] & 7) << 1) | (extra[1] >> 7), cc = (extra[1] >> 3) & 15;
            static const int rates[13] = {96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000, 12000, 11025, 8000, 7350};
            if (aot == 2 && fi < 13 && cc >= 1 && cc <= 7) {
                copy = true;
                out->present = true; out->copied = true;
                out->sample_rate = rates[fi]; out->channels = cc == 7 ? 8 : cc;
                out->frame_samples = AAC_FRAME; out->priming = 0; out->bitrate = bitrate;
                out->asc.assign(extra, extra + extra_size);
                return VCPENC_OK;
            }
        }
        if (!L->audio_ok) { set_err(err, errlen, "input has an audio stream and the loaded FFmpeg libraries lack the calls to transcode it (pass -an to drop it)"); return VCPENC_E_AUDIO; }
        void* codec = L->avcodec_find_decoder(codec_id);
        if (!codec) { set_err(err, errlen, "no decoder for the input's audio (codec id %d)", codec_id); return VCPENC_E_AUDIO; }
        dec = L->avcodec_alloc_context3(codec);
        if (!dec || L->avcodec_parameters_to_context(dec, par) < 0 || L->avcodec_open2(dec, codec, nullptr) < 0) { set_err(err, errlen, "cannot open the audio decoder"); return VCPENC_E_AUDIO; }
        dframe = L->av_frame_alloc();
        opkt = L->av_packet_alloc();
        out->present = true; out->copied = false; out->bitrate = bitrate;
        out->frame_samples = AAC_FRAME; out->priming = AAC_FRAME;   // the native encoder's delay: one frame
        return VCPENC_OK;
    }

    // first decoded frame: layout / rate / format are known -> encoder and converter
    int configure(char* err, size_t errlen) {
        const AVFrameHead* f = static_cast<const AVFrameHead*>(dframe);
        in_fmt = f->format;
        int64_t v = 0;
        if (L->av_opt_get_int(dec, "ar", 0, &v) < 0 || v < 4000 || v > 768000) { set_err(err, errlen, "audio: cannot read the decoder's sample rate"); return VCPENC_E_AUDIO; }
        in_rate = (int)v;
        AVChannelLayoutPub inl{};
        if (L->av_opt_get_chlayout(dec, "ch_layout", 0, &inl) < 0 || inl.nb_channels < 1 || inl.nb_channels > 8) { set_err(err, errlen, "audio: unsupported channel layout"); return VCPENC_E_AUDIO; }
        if (inl.order != 1) L->av_channel_layout_default(&inl, inl.nb_channels);   // unspecified order: the default layout of that many channels
        const int out_rate = aac_rate_index(in_rate) >= 0 ? in_rate : 48000;
        void* codec = L->avcodec_find_encoder_by_name("aac");
        if (!codec) { set_err(err, errlen, "the loaded libavcodec has no aac encoder"); return VCPENC_E_AUDIO; }
        AVChannelLayoutPub outl = inl;
        for (int attempt = 0; attempt < 2; attempt++) {
            if (attempt == 1) { if (inl.nb_channels <= 2) break; L->av_channel_layout_default(&outl, 2); }   // layouts the encoder refuses are mixed down to stereo
            enc = L->avcodec_alloc_context3(codec);
            if (!enc) break;
            const AVOptionHead* o_ar = static_cast<const AVOptionHead*>(L->av_opt_find(enc, "ar", nullptr, 0, 0));
            const AVOptionHead* o_cl = static_cast<const AVOptionHead*>(L->av_opt_find(enc, "ch_layout", nullptr, 0, 0));
            const AVOptionHead* o_ac = static_cast<const AVOptionHead*>(L->av_opt_find(enc, "ac", nullptr, 0, 0));
            int fmt_off = -1;
            if (o_ar && o_cl && o_cl->offset == o_ar->offset + 8) fmt_off = o_ar->offset + 4;             // { sample_rate; sample_fmt; ch_layout } (libavcodec 61+)
            else if (o_ar && o_ac && o_ac->offset == o_ar->offset + 4) fmt_off = o_ar->offset + 8;        // { sample_rate; channels; sample_fmt } (59, 60)
            if (fmt_off < 0) { set_err(err, errlen, "audio: unknown AVCodecContext layout (libavcodec ABI)"); return VCPENC_E_AUDIO; }
            L->av_opt_set_int(enc, "b", bitrate, 0);
            L->av_opt_set_int(enc, "ar", out_rate, 0);
            L->av_opt_set_chlayout(enc, "ch_layout", &outl, 0);
            *reinterpret_cast<int*>(static_cast<uint8_t*>(enc) + fmt_off) = SAMPLE_FMT_FLTP;
            if (L->avcodec_open2(enc, codec, nullptr) >= 0) break;
            L->avcodec_free_context(&enc);
            enc = nullptr;
        }
        if (!enc) { set_err(err, errlen, "cannot open the aac encoder (%d Hz, %d channels)", out_rate, inl.nb_channels); return VCPENC_E_AUDIO; }
        if (L->av_opt_get_int(enc, "frame_size", 0, &v) < 0 || v != AAC_FRAME) { set_err(err, errlen, "audio: unexpected aac frame size"); return VCPENC_E_AUDIO; }
        if (L->swr_alloc_set_opts2(&swr, &outl, SAMPLE_FMT_FLTP, out_rate, &inl, in_fmt, in_rate, 0, nullptr) < 0 || L->swr_init(swr) < 0) {
            set_err(err, errlen, "audio: cannot set up the sample converter"); return VCPENC_E_AUDIO;
        }
        carrier = L->av_frame_alloc();
        if (!carrier || L->av_frame_ref(carrier, dframe) < 0) { set_err(err, errlen, "audio: cannot reference a decoded frame"); return VCPENC_E_AUDIO; }
        fifo.assign((size_t)outl.nb_channels, std::vector<float>());
        {
            std::lock_guard<std::mutex> lk(*mu);
            out->sample_rate = out_rate; out->channels = outl.nb_channels;
            const int fi = aac_rate_index(out_rate), cc = outl.nb_channels == 8 ? 7 : outl.nb_channels;
            out->asc = {(uint8_t)((2 << 3) | (fi >> 1)), (uint8_t)(((fi & 1) << 7) | (cc << 3))};   // AAC-LC
        }
        configured = true;
        return VCPENC_OK;
    }

    int drain_encoder() {
        for (;;) {
            const int r = L->avcodec_receive_packet(enc, opkt);
            if (r < 0) return (r == -11 || r == AVERROR_EOF_) ? 0 : r;
            const AVPacketHead* ph = static_cast<const AVPacketHead*>(opkt);
            if (ph->size > 0) emit(ph->data, ph->size);
            L->av_packet_unref(opkt);
        }
    }

    // hand `n` samples per channel (n <= AAC_FRAME) from the front of the FIFO to the encoder
    int encode_block(int n, char* err, size_t errlen) {
        AVFrameHead* c = static_cast<AVFrameHead*>(carrier);
        const size_t nch = fifo.size();
        scratch.assign(nch * AAC_FRAME, 0.0f);
        for (size_t ch = 0; ch < nch; ch++) {
            memcpy(&scratch[ch * AAC_FRAME], fifo[ch].data(), (size_t)n * sizeof(float));
            fifo[ch].erase(fifo[ch].begin(), fifo[ch].begin() + n);
            c->data[ch] = reinterpret_cast<uint8_t*>(&scratch[ch * AAC_FRAME]);
        }
        for (size_t ch = nch; ch < 8; ch++) c->data[ch] = nullptr;
        c->linesize[0] = AAC_FRAME * (int)sizeof(float);
        c->extended_data = c->data;
        c->nb_samples = AAC_FRAME;     // a short last block is padded with silence
        c->format = SAMPLE_FMT_FLTP;
        *reinterpret_cast<int64_t*>(reinterpret_cast<uint8_t*>(carrier) + 136) = sent;   // pts (same offset in every libavutil since 52)
        sent += AAC_FRAME;
        const int r = L->avcodec_send_frame(enc, carrier);
        if (r < 0 && r != -11) { set_err(err, errlen, "aac encoder rejected a frame (%d)", r); return VCPENC_E_AUDIO; }
        if (drain_encoder() < 0) { set_err(err, errlen, "aac encoder failed"); return VCPENC_E_AUDIO; }
        return VCPENC_OK;
    }

    int take_decoded(char* err, size_t errlen) {
        for (;;) {
            const int r = L->avcodec_receive_frame(dec, dframe);
            if (r < 0) return VCPENC_OK;   // needs more input / end
            if (!configured) { const int rc = configure(err, errlen); if (rc) return rc; }
            const AVFrameHead* f = static_cast<const AVFrameHead*>(dframe);
            if (f->format != in_fmt) { set_err(err, errlen, "audio: sample format changes mid-stream are not supported"); return VCPENC_E_AUDIO; }
            const size_t nch = fifo.size();
            int cap = L->swr_get_out_samples(swr, f->nb_samples);
            if (cap < f->nb_samples) cap = f->nb_samples + 64;
            std::vector<float> tmp(nch * (size_t)cap);
            uint8_t* op[8] = {nullptr};
            for (size_t ch = 0; ch < nch; ch++) op[ch] = reinterpret_cast<uint8_t*>(&tmp[ch * (size_t)cap]);
            const int got = L->swr_convert(swr, op, cap, const_cast<const uint8_t**>(f->extended_data), f->nb_samples);
            L->av_frame_unref(dframe);
            if (got < 0) { set_err(err, errlen, "audio: sample conversion failed"); return VCPENC_E_AUDIO; }
            for (size_t ch = 0; ch < nch; ch++) fifo[ch].insert(fifo[ch].end(), &tmp[ch * (size_t)cap], &tmp[ch * (size_t)cap] + got);
            while (fifo[0].size() >= (size_t)AAC_FRAME) { const int rc = encode_block(AAC_FRAME, err, errlen); if (rc) return rc; }
        }
    }

    int packet(void* pkt, char* err, size_t errlen) {
        const AVPacketHead* ph = static_cast<const AVPacketHead*>(pkt);
        if (copy) { if (ph->size > 0) emit(ph->data, ph->size); return VCPENC_OK; }
        const int s = L->avcodec_send_packet(dec, pkt);
        if (s < 0 && s != -11) return VCPENC_OK;   // a damaged audio packet is skipped, as ffmpeg does
        return take_decoded(err, errlen);
    }

    int finish(char* err, size_t errlen) {
        if (copy || !dec) return VCPENC_OK;
        L->avcodec_send_packet(dec, nullptr);
        int rc = take_decoded(err, errlen);
        if (rc || !configured) return rc;
        // what the converter still holds, then the partial last block
        {
            const size_t nch = fifo.size();
            const int cap = L->swr_get_out_samples(swr, 0) + 64;
            std::vector<float> tmp(nch * (size_t)cap);
            uint8_t* op[8] = {nullptr};
            for (size_t ch = 0; ch < nch; ch++) op[ch] = reinterpret_cast<uint8_t*>(&tmp[ch * (size_t)cap]);
            const int got = L->swr_convert(swr, op, cap, nullptr, 0);
            for (size_t ch = 0; got > 0 && ch < nch; ch++) fifo[ch].insert(fifo[ch].end(), &tmp[ch * (size_t)cap], &tmp[ch * (size_t)cap] + got);
        }
        while (!fifo[0].empty()) { rc = encode_block((int)std::min<size_t>(fifo[0].size(), AAC_FRAME), err, errlen); if (rc) return rc; }
        L->avcodec_send_frame(enc, nullptr);
        if (drain_encoder() < 0) { set_err(err, errlen, "aac encoder failed while flushing"); return VCPENC_E_AUDIO; }
        return VCPENC_OK;
    }
};

struct LavSource : FrameSource {
    Lav* L = nullptr;
    void* fmtctx = nullptr;
    void* dec = nullptr;
    void* pkt = nullptr;
    void* frame = nullptr;
    int vstream = -1, astream = -1;
    bool draining = false, done = false;
    std::unique_ptr<LavAudio> aud;
    int audio_rc = 0;

    ~LavSource() override {
        aud.reset();
        if (!L) return;
        if (frame) L->av_frame_free(&frame);
        if (pkt) L->av_packet_free(&pkt);
        if (dec) L->avcodec_free_context(&dec);
        if (fmtctx) L->avformat_close_input(&fmtctx);
    }

    int open(const char* path, bool drop_audio, int audio_bitrate, char* err, size_t errlen) {
        L = load_lav();
        if (!L->ok) { set_err(err, errlen, "container input needs FFmpeg's demuxer/decoder: %s", L->why.c_str()); return VCPENC_E_FORMAT; }
        if (L->avformat_open_input(&fmtctx, path, nullptr, nullptr) < 0) { fmtctx = nullptr; set_err(err, errlen, "cannot open/parse %s", path); return VCPENC_E_FORMAT; }
        if (L->avformat_find_stream_info(fmtctx, nullptr) < 0) { set_err(err, errlen, "no stream information in %s", path); return VCPENC_E_FORMAT; }
        vstream = L->av_find_best_stream(fmtctx, AVMEDIA_TYPE_VIDEO, -1, -1, nullptr, 0);
        if (vstream < 0) { set_err(err, errlen, "no video stream in %s", path); return VCPENC_E_FORMAT; }
        astream = drop_audio ? -1 : L->av_find_best_stream(fmtctx, AVMEDIA_TYPE_AUDIO, -1, vstream, nullptr, 0);
        // AVFormatContext { av_class, iformat, oformat, priv_data, pb; int ctx_flags; unsigned nb_streams; AVStream** streams }
        const uint8_t* fc = static_cast<const uint8_t*>(fmtctx);
        const unsigned nb_streams = *reinterpret_cast<const unsigned*>(fc + 44);
        if ((unsigned)vstream >= nb_streams || nb_streams > 4096 || (astream >= 0 && (unsigned)astream >= nb_streams)) { set_err(err, errlen, "unexpected AVFormatContext layout (libavformat ABI)"); return VCPENC_E_INTERNAL; }
        void** streams = *reinterpret_cast<void** const*>(fc + 48);
        if (astream >= 0) {
            void* ast = streams[astream];
            if (*reinterpret_cast<const int*>(static_cast<const uint8_t*>(ast) + 8) != astream) { set_err(err, errlen, "unexpected AVStream layout (libavformat ABI)"); return VCPENC_E_INTERNAL; }
            void* apar = *reinterpret_cast<void* const*>(static_cast<const uint8_t*>(ast) + 16);
            if (static_cast<const int*>(apar)[0] != AVMEDIA_TYPE_AUDIO) { set_err(err, errlen, "unexpected AVCodecParameters layout (libavcodec ABI)"); return VCPENC_E_INTERNAL; }
            aud = std::make_unique<LavAudio>();
            const int arc = aud->open(L, apar, astream, audio_bitrate, &audio, &audio_mu, err, errlen);
            if (arc) return arc;
        }
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
        // Expected number of pictures (sizes the page-locked buffers and the moov reserve): container duration x frame rate.
        // AVFormatContext keeps { int64 start_time; int64 duration; int64 bit_rate; unsigned packet_size; ... } together in every
        // libavformat since 3.x; "packetsize" has an AVOption, which gives the anchor without headers.
        if (L->av_opt_find) {
            const AVOptionHead* o = static_cast<const AVOptionHead*>(L->av_opt_find(fmtctx, "packetsize", nullptr, 0, 0));
            if (o && o->offset >= 24 && o->offset < 2048) {
                const int64_t dur = *reinterpret_cast<const int64_t*>(fc + o->offset - 16);   // AV_TIME_BASE units (microseconds)
                if (dur > 0 && dur < (int64_t)48 * 3600 * 1000000) est_frames = (long)(dur * fps_num / ((int64_t)fps_den * 1000000)) + 2 + (long)(dur * fps_num / ((int64_t)fps_den * 1000000)) / 50;
            }
        }
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
                } else if (aud && ph->stream_index == astream) {
                    const int arc = aud->packet(pkt, err, errlen);
                    if (arc) { L->av_packet_unref(pkt); return -arc; }
                }
                L->av_packet_unref(pkt);
            }
        }
    }

    int finish_audio(char* err, size_t errlen) override {
        if (!aud) return 0;
        // audio packets behind the last video packet
        while (!draining && L->av_read_frame(fmtctx, pkt) >= 0) {
            const AVPacketHead* ph = static_cast<const AVPacketHead*>(pkt);
            int arc = 0;
            if (ph->stream_index == astream) arc = aud->packet(pkt, err, errlen);
            L->av_packet_unref(pkt);
            if (arc) return arc;
        }
        return aud->finish(err, errlen);
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

int open_container_source(const char* path, bool drop_audio, int audio_bitrate, std::unique_ptr<FrameSource>* out, char* err, size_t errlen) {
    auto s = std::make_unique<LavSource>();
    const int rc = s->open(path, drop_audio, audio_bitrate, err, errlen);
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
    int rc = vcp::open_container_source(path, true, 0, &src, err, errlen);
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

// Host-only: run the AUDIO side of a container input the way vcpenc_transcode does (stream copy of AAC-LC, else
// decode -> planar float -> libavcodec `aac` at `audio_bitrate`) and return the access units.  The pictures are decoded
// and dropped.  *copied = 1 for a stream copy.  Returns VCPENC_E_FORMAT when the input has no audio stream.
extern "C" int vcpenc_probe_audio(const char* path, int audio_bitrate, int* sample_rate, int* channels, int* copied, int* priming,
                                  uint8_t* asc, int asc_cap, int* asc_len, uint8_t* data, size_t data_cap, size_t* data_len,
                                  uint32_t* sizes, int sizes_cap, int* nframes, char* err, size_t errlen) {
    if (!path) { vcp::set_err(err, errlen, "bad arguments"); return VCPENC_E_ARGS; }
    std::unique_ptr<vcp::FrameSource> src;
    int rc = vcp::open_container_source(path, false, audio_bitrate, &src, err, errlen);
    if (rc) return rc;
    if (!src->audio.present) { vcp::set_err(err, errlen, "no audio stream in %s", path); return VCPENC_E_FORMAT; }
    std::vector<uint8_t> scratch(src->fbytes() * 4);
    for (;;) {
        const int n = src->read(scratch.data(), 4, err, errlen);
        if (n < 0) return -n;
        if (n < 4) break;
    }
    rc = src->finish_audio(err, errlen);
    if (rc) return rc;
    const vcp::AudioTrack& a = src->audio;
    if (sample_rate) *sample_rate = a.sample_rate;
    if (channels) *channels = a.channels;
    if (copied) *copied = a.copied ? 1 : 0;
    if (priming) *priming = a.priming;
    if (asc_len) *asc_len = (int)a.asc.size();
    if (asc && asc_cap >= (int)a.asc.size()) memcpy(asc, a.asc.data(), a.asc.size());
    if (nframes) *nframes = (int)a.sizes.size();
    if (data_len) *data_len = a.data.size();
    if (data && sizes) {
        if (a.data.size() > data_cap || (int)a.sizes.size() > sizes_cap) { vcp::set_err(err, errlen, "audio buffers too small"); return VCPENC_E_OVERFLOW; }
        memcpy(data, a.data.data(), a.data.size());
        memcpy(sizes, a.sizes.data(), a.sizes.size() * sizeof(uint32_t));
    }
    return VCPENC_OK;
}
