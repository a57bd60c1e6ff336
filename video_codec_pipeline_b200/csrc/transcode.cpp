// vcpenc_transcode — the drop-in for runFFmpegWithTimeout
// (/root/reference/cmd/consumer.go:370-394): input path, output path, whitespace-split option
// tokens, a timeout and a cancel flag in; a complete output file (MP4, or raw .h264) or an
// error class out.  Semantics kept from the reference:
//   * `-y`: an existing output is overwritten (:376)
//   * nil error <=> the file is complete on return (:386,393)
//   * deadline  -> "编码超时", parent cancel -> "任务被取消" (:387-392)
//   * on failure the caller removes the partial output (:264); we also do not leave one
//   * stdout/stderr are inherited: progress lines go to stderr in key=value form
// Frames are pulled in chunks of whole GOPs, pushed through the device session (K1..K5) and
// the resulting access units are appended to the muxer; no frame ever takes a CPU encode path.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "frontend.h"
#include "host_bits.h"
#include "host_util.h"
#include "vcp_algo.h"
#include <sys/stat.h>

using namespace vcp;

namespace {

size_t fbytes(int w, int h) { return (size_t)w * h + 2 * (size_t)((w + 1) / 2) * ((h + 1) / 2); }

struct Y4mSource : FrameSource {
    FILE* f = nullptr;
    ~Y4mSource() override { if (f) fclose(f); }
    int open(const char* path, char* err, size_t errlen) {
        f = fopen(path, "rb");
        if (!f) { set_err(err, errlen, "cannot open %s", path); return VCPENC_E_IO; }
        char line[512];
        if (!fgets(line, sizeof line, f) || strncmp(line, "YUV4MPEG2", 9)) { set_err(err, errlen, "not a YUV4MPEG2 stream"); return VCPENC_E_FORMAT; }
        fps_num = 25; fps_den = 1;
        std::string cs = "420";
        for (char* t = strtok(line + 9, " \n"); t; t = strtok(nullptr, " \n")) {
            if (t[0] == 'W') width = atoi(t + 1);
            else if (t[0] == 'H') height = atoi(t + 1);
            else if (t[0] == 'F') { if (sscanf(t + 1, "%d:%d", &fps_num, &fps_den) != 2) { fps_num = 25; fps_den = 1; } }
            else if (t[0] == 'C') cs = t + 1;
            else if (t[0] == 'I' && t[1] != 'p' && t[1] != '?') { set_err(err, errlen, "interlaced y4m not supported"); return VCPENC_E_FORMAT; }
        }
        const bool deep = cs.find("p10") != std::string::npos || cs.find("p12") != std::string::npos || cs.find("p16") != std::string::npos;
        if (!deep && cs.compare(0, 3, "420") == 0) fmt = VCPENC_FMT_YUV420P;
        else if (!deep && cs == "422") fmt = VCPENC_FMT_YUV422P;
        else if (!deep && cs == "444") fmt = VCPENC_FMT_YUV444P;
        else {
            set_err(err, errlen, "y4m colourspace C%s not supported (8-bit 4:2:0 / 4:2:2 / 4:4:4 only)", cs.c_str());
            return VCPENC_E_FORMAT;
        }
        if (width <= 0 || height <= 0) { set_err(err, errlen, "bad y4m header"); return VCPENC_E_FORMAT; }
        return VCPENC_OK;
    }
    int read(uint8_t* dst, int max, char* err, size_t errlen) override {
        const size_t fb = fbytes();
        int n = 0;
        while (n < max) {
            char line[128];
            if (!fgets(line, sizeof line, f)) break;
            if (strncmp(line, "FRAME", 5)) { set_err(err, errlen, "y4m: missing FRAME marker"); return -VCPENC_E_FORMAT; }
            if (fread(dst + (size_t)n * fb, 1, fb, f) != fb) { set_err(err, errlen, "y4m: truncated frame"); return -VCPENC_E_FORMAT; }
            n++;
        }
        return n;
    }
};

struct RawSource : FrameSource {
    FILE* f = nullptr;
    ~RawSource() override { if (f) fclose(f); }
    int open(const char* path, const vcpenc_params& p, int pixfmt, char* err, size_t errlen) {
        if (p.in_width <= 0 || p.in_height <= 0) { set_err(err, errlen, "raw input needs -s WxH"); return VCPENC_E_FORMAT; }
        fmt = pixfmt;
        width = p.in_width; height = p.in_height; fps_num = p.fps_num; fps_den = p.fps_den;
        f = fopen(path, "rb");
        if (!f) { set_err(err, errlen, "cannot open %s", path); return VCPENC_E_IO; }
        return VCPENC_OK;
    }
    int read(uint8_t* dst, int max, char*, size_t) override {
        const size_t fb = fbytes();
        int n = 0;
        while (n < max && fread(dst + (size_t)n * fb, 1, fb, f) == fb) n++;
        return n;
    }
};

bool ends_with(const std::string& s, const char* suf) {
    const size_t n = strlen(suf);
    if (s.size() < n) return false;
    for (size_t i = 0; i < n; i++)
        if (tolower((unsigned char)s[s.size() - n + i]) != suf[i]) return false;
    return true;
}

struct PinnedBuf {
    uint8_t* p = nullptr;
    ~PinnedBuf() { if (p) vcpenc_host_free(p); }
};

}  // namespace

static thread_local int t_device = 0;

// Per-thread reuse across tasks: the session (all device buffers) and the pinned staging buffer.
// Measured on the task-flow harness: session create + destroy 0.1-1.4 s and cudaMallocHost 0.1-0.3 s
// per task against 0.13 s of GPU work for a 120-frame 1080p clip.
namespace {
constexpr int kCacheSlots = 16;   // mixed task sizes (720p / 1080p / 4K of config #5) each keep their session
struct CachedSession {
    vcpenc_session* ses = nullptr;
    vcpenc_params key{};
    int max_frames = 0, device = -1;
    unsigned long stamp = 0;
};
struct ThreadCache {
    CachedSession slot[kCacheSlots];
    unsigned long clock = 0;
    uint8_t* pinned = nullptr;
    size_t pinned_bytes = 0;
};
thread_local ThreadCache t_cache;
bool cache_enabled() { static const bool on = getenv("VCPENC_NO_CACHE") == nullptr; return on; }
bool same_key(vcpenc_params a, vcpenc_params b) { a.first_gop = b.first_gop = 0; return memcmp(&a, &b, sizeof a) == 0; }
}  // namespace

// session of `want` frames on `device` for parameters p: from the calling thread's cache, else new
// (and cached).  *cached tells whether the cache owns it.
int acquire_session(const vcpenc_params& p, int device, int want, vcpenc_session** out, bool* cached, char* err, size_t errlen) {
    *cached = false;
    if (cache_enabled())
        for (auto& c : t_cache.slot)
            if (c.ses && c.device == device && c.max_frames >= want && same_key(c.key, p)) {
                c.stamp = ++t_cache.clock; *out = c.ses; *cached = true;
                return VCPENC_OK;
            }
    CachedSession* victim = nullptr;
    if (cache_enabled()) {   // an empty slot, else the one with the same parameters but too small, else the least recently used
        for (auto& c : t_cache.slot) if (!c.ses) { victim = &c; break; }
        if (!victim) for (auto& c : t_cache.slot) if (same_key(c.key, p) && c.device == device) { victim = &c; break; }
        if (!victim) { victim = &t_cache.slot[0]; for (auto& c : t_cache.slot) if (c.stamp < victim->stamp) victim = &c; }
        if (victim->ses) { vcpenc_session_destroy(victim->ses); *victim = CachedSession(); }
    }
    const int rc = vcpenc_session_create(&p, device, want, out, err, errlen);
    if (rc) return rc;
    if (victim) { victim->ses = *out; victim->key = p; victim->max_frames = want; victim->device = device; victim->stamp = ++t_cache.clock; *cached = true; }
    return VCPENC_OK;
}
void forget_session(vcpenc_session* ses) {   // after a failure the session's state is not trusted
    vcpenc_session_destroy(ses);
    for (auto& c : t_cache.slot) if (c.ses == ses) c = CachedSession();
}

// GPUs one task may use: VCPENC_GPUS=N|all shards the closed GOPs of every chunk across N devices
// (the calling thread's device first), concatenated on the host in GOP order -- no collective.
int task_gpus() {
    const char* e = getenv("VCPENC_GPUS");
    if (!e) return 1;
    const int have = vcpenc_device_count();
    const int want = !strcmp(e, "all") ? have : atoi(e);
    return std::max(1, std::min(want, have));
}

extern "C" void vcpenc_thread_release(void) {
    for (auto& c : t_cache.slot) if (c.ses) vcpenc_session_destroy(c.ses);
    if (t_cache.pinned) vcpenc_host_free(t_cache.pinned);
    t_cache = ThreadCache();
}
extern "C" int vcpenc_set_thread_device(int device) {
    if (device < 0 || device >= vcpenc_device_count()) return VCPENC_E_NODEVICE;
    t_device = device;
    return VCPENC_OK;
}

extern "C" int vcpenc_transcode(const char* input, const char* output, int argc, const char* const* argv,
                                int timeout_ms, volatile int* cancel, char* err, size_t errlen) {
    using clock = std::chrono::steady_clock;
    const auto t0 = clock::now();
    // VCPENC_TRACE=1: where the wall time of a task goes (open / pinned alloc / read+decode / session / GPU / mux)
    const bool trace = getenv("VCPENC_TRACE") != nullptr;
    double t_open = 0, t_alloc = 0, t_read = 0, t_create = 0, t_gpu = 0, t_mux = 0;
    auto lap = [&](clock::time_point& from) { const auto now = clock::now(); const double d = std::chrono::duration<double>(now - from).count(); from = now; return d; };
    auto tl = t0;
    if (!input || !output) { set_err(err, errlen, "bad arguments"); return VCPENC_E_ARGS; }
    vcpenc_params p;
    int rc = vcpenc_parse_args(argc, argv, &p, err, errlen);
    if (rc) return rc;
    if (vcpenc_device_count() <= 0) { set_err(err, errlen, "no CUDA device available (libvcpenc has no CPU fallback)"); return VCPENC_E_NODEVICE; }

    const std::string in = input, outp = output;
    std::unique_ptr<FrameSource> src;
    if (ends_with(in, ".y4m")) {
        auto s = std::make_unique<Y4mSource>();
        rc = s->open(input, err, errlen);
        if (rc) return rc;
        src = std::move(s);
    } else if (ends_with(in, ".yuv") || ends_with(in, ".nv12") || ends_with(in, ".rgb") || ends_with(in, ".bgr")) {
        // headerless frames: the size comes from -s WxH, the pixel format from the extension
        const int pf = ends_with(in, ".nv12") ? VCPENC_FMT_NV12 : ends_with(in, ".rgb") ? VCPENC_FMT_RGB24
                     : ends_with(in, ".bgr") ? VCPENC_FMT_BGR24 : VCPENC_FMT_YUV420P;
        auto s = std::make_unique<RawSource>();
        rc = s->open(input, p, pf, err, errlen);
        if (rc) return rc;
        src = std::move(s);
        p.in_width = p.in_height = 0;   // consumed as the input size
    } else {
        // container input: what the producer forwards (.mp4 .mkv .avi .mov .webm, cmd/producer.go:485-488)
        rc = open_container_source(input, p.drop_audio != 0, &src, err, errlen);
        if (rc) return rc;
        if (p.in_width > 0 && p.width == 0) { p.width = p.in_width; p.height = p.in_height; }   // -s after -i = output size
    }
    // output size: -vf scale=W:H (negative = keep aspect, rounded to even), else -s WxH on a
    // self-describing input, else the input size
    {
        int ow = p.width, oh = p.height;
        if (ow == 0 && oh == 0 && p.in_width > 0) { ow = p.in_width; oh = p.in_height; }
        if (ow == 0 && oh == 0) { ow = src->width; oh = src->height; }
        if (ow < 0) ow = (int)(((long long)src->width * oh / src->height + 1) & ~1LL);
        if (oh < 0) oh = (int)(((long long)src->height * ow / src->width + 1) & ~1LL);
        p.width = ow; p.height = oh;
    }
    t_open = lap(tl);
    p.in_fmt = src->fmt;
    p.in_width = src->width; p.in_height = src->height;
    p.fps_num = src->fps_num; p.fps_den = src->fps_den;
    if ((p.width & 1) || (p.height & 1) || p.width < 16 || p.height < 16) { set_err(err, errlen, "unsupported picture size %dx%d", p.width, p.height); return VCPENC_E_FORMAT; }
    if (p.slices == 0) p.slices = vcp_auto_slices((p.height + 15) / 16, p.entropy);
    if (p.slices > (p.height + 15) / 16) p.slices = (p.height + 15) / 16;

    const size_t fb = std::max(fbytes(p.width, p.height), src->fbytes());
    // chunk: whole GOPs, at most ~3 GiB of raw frames resident per pass
    size_t chunk_bytes = ((size_t)3 << 30) * (size_t)task_gpus();
    if (const char* e = getenv("VCPENC_CHUNK_BYTES")) { const long long v = atoll(e); if (v > 0) chunk_bytes = (size_t)v; }   // tests: force many chunks
    int chunk = (int)std::max<size_t>(1, chunk_bytes / fb);
    {   // short clips: do not page-lock more host memory than the file can fill
        struct stat sb;
        if (stat(input, &sb) == 0 && sb.st_size > 0)
            chunk = (int)std::min<size_t>((size_t)chunk, (size_t)sb.st_size / src->fbytes() + 1);
    }
    chunk = std::max(p.gop, (chunk + p.gop - 1) / p.gop * p.gop);
    PinnedBuf frames;   // owns the buffer only when the per-thread cache is off
    uint8_t* fbuf = nullptr;
    if (cache_enabled()) {
        if (t_cache.pinned_bytes < (size_t)chunk * fb) {
            if (t_cache.pinned) vcpenc_host_free(t_cache.pinned);
            t_cache.pinned = (uint8_t*)vcpenc_host_alloc((size_t)chunk * fb);
            t_cache.pinned_bytes = t_cache.pinned ? (size_t)chunk * fb : 0;
        }
        fbuf = t_cache.pinned;
    } else {
        frames.p = (uint8_t*)vcpenc_host_alloc((size_t)chunk * fb);
        fbuf = frames.p;
    }
    if (!fbuf) { set_err(err, errlen, "cannot allocate %zu bytes of pinned host memory", (size_t)chunk * fb); return VCPENC_E_CUDA; }
    t_alloc = lap(tl);

    const int ndev = task_gpus();
    struct Shard {
        int device = 0; vcpenc_session* ses = nullptr; bool cached = false;
        std::vector<uint8_t> bits; std::vector<vcpenc_frame_info> info;
        int f0 = 0, n = 0, g0 = 0; size_t len = 0; int rc = 0; char err[256] = {0};
    };
    std::vector<Shard> shards((size_t)ndev);
    for (int d = 0; d < ndev; d++) shards[d].device = (t_device + d) % std::max(1, vcpenc_device_count());
    std::vector<uint8_t> mdat, annexb_all;
    ParamSets psets;
    std::vector<Mp4Sample> samples;
    const bool raw_out = ends_with(outp, ".h264") || ends_with(outp, ".264") || ends_with(outp, ".h265") || ends_with(outp, ".265") || ends_with(outp, ".hevc");
    long total = 0;
    int gop_index = 0;
    auto fail = [&](int code) {
        for (auto& sh : shards) if (sh.ses) { forget_session(sh.ses); sh.ses = nullptr; }
        remove(output);
        return code;
    };
    for (;;) {
        if (cancel && *cancel) { set_err(err, errlen, "任务被取消"); return fail(VCPENC_E_CANCELLED); }
        if (timeout_ms > 0 && std::chrono::duration_cast<std::chrono::milliseconds>(clock::now() - t0).count() > timeout_ms) {
            set_err(err, errlen, "编码超时 (>%dms)", timeout_ms);
            return fail(VCPENC_E_TIMEOUT);
        }
        const int n = src->read(fbuf, chunk, err, errlen);
        t_read += lap(tl);
        if (n < 0) return fail(-n);
        if (n == 0) break;
        // closed GOPs of this chunk, contiguous ranges per device
        const int ngops = (n + p.gop - 1) / p.gop;
        const int used = std::min(ndev, ngops);
        for (int d = 0; d < ndev; d++) {
            Shard& sh = shards[d];
            const int ga = d < used ? (int)((long long)ngops * d / used) : 0, gb = d < used ? (int)((long long)ngops * (d + 1) / used) : 0;
            sh.f0 = ga * p.gop; sh.n = std::min(n, gb * p.gop) - sh.f0; sh.g0 = gop_index + ga; sh.rc = 0; sh.len = 0;
            if (sh.n <= 0) { sh.n = 0; continue; }
            if (!sh.ses) {
                p.first_gop = 0;
                const int want = std::min((chunk / p.gop + used - 1) / used * p.gop, std::max(sh.n, 1));
                rc = acquire_session(p, sh.device, std::max(want, sh.n), &sh.ses, &sh.cached, err, errlen);
                if (rc) return fail(rc);
            }
            if (sh.bits.size() < (size_t)sh.n * fb / 2 + (1 << 20)) sh.bits.resize((size_t)sh.n * fb / 2 + (1 << 20));
            if (sh.info.size() < (size_t)sh.n) sh.info.resize((size_t)sh.n);
        }
        t_create += lap(tl);
        auto run = [&](Shard& sh) {
            if (!sh.n) return;
            vcpenc_session_set_first_gop(sh.ses, sh.g0);          // idr_pic_id parity continues across ranges and chunks
            // streamed: the pinned chunk buffer outlives the encode, whose GOP groups start as their frames land
            sh.rc = vcpenc_session_upload_async(sh.ses, fbuf + (size_t)sh.f0 * src->fbytes(), sh.n, sh.err, sizeof sh.err);
            if (!sh.rc) sh.rc = vcpenc_session_encode(sh.ses, nullptr, sh.err, sizeof sh.err);
            if (!sh.rc) {
                sh.rc = vcpenc_session_download(sh.ses, sh.bits.data(), sh.bits.size(), &sh.len, sh.info.data(), nullptr, sh.err, sizeof sh.err);
                if (sh.rc == VCPENC_E_OVERFLOW) {
                    sh.bits.resize((size_t)sh.n * fb + (1 << 20));
                    sh.rc = vcpenc_session_download(sh.ses, sh.bits.data(), sh.bits.size(), &sh.len, sh.info.data(), nullptr, sh.err, sizeof sh.err);
                }
            }
        };
        if (used <= 1) run(shards[0]);
        else {
            std::vector<std::thread> th;
            for (int d = 0; d < used; d++) th.emplace_back(run, std::ref(shards[d]));
            for (auto& t : th) t.join();
        }
        for (auto& sh : shards) if (sh.n && sh.rc) { set_err(err, errlen, "%s", sh.err); return fail(sh.rc); }
        t_gpu += lap(tl);
        for (auto& sh : shards) {            // host concatenation in GOP order
            if (!sh.n) continue;
            if (raw_out) { annexb_all.insert(annexb_all.end(), sh.bits.begin(), sh.bits.begin() + sh.len); continue; }
            for (int i = 0; i < sh.n; i++) {
                Mp4Sample sm{mdat.size(), 0, sh.info[i].is_idr != 0};
                for (const auto& nal : split_annexb(sh.bits.data() + sh.info[i].offset, sh.info[i].size)) {
                    if (!nal.n) continue;
                    if (psets.take(p.codec, nal)) continue;
                    const uint32_t k = (uint32_t)nal.n;
                    const uint8_t h[4] = {(uint8_t)(k >> 24), (uint8_t)(k >> 16), (uint8_t)(k >> 8), (uint8_t)k};
                    mdat.insert(mdat.end(), h, h + 4);
                    mdat.insert(mdat.end(), nal.p, nal.p + nal.n);
                }
                sm.size = (uint32_t)(mdat.size() - sm.offset);
                samples.push_back(sm);
            }
        }
        total += n;
        gop_index += ngops;
        if (n < chunk) break;
    }
    for (auto& sh : shards) { if (sh.ses && !sh.cached) vcpenc_session_destroy(sh.ses); sh.ses = nullptr; }
    if (total == 0) { set_err(err, errlen, "input has no frames"); remove(output); return VCPENC_E_FORMAT; }
    if (raw_out) {
        FILE* f = fopen(output, "wb");
        if (!f) { set_err(err, errlen, "cannot create %s", output); return VCPENC_E_IO; }
        const bool ok = fwrite(annexb_all.data(), 1, annexb_all.size(), f) == annexb_all.size();
        if (fclose(f) != 0 || !ok) { set_err(err, errlen, "short write to %s", output); remove(output); return VCPENC_E_IO; }
    } else {
        rc = write_mp4(p, psets, samples, mdat.data(), mdat.size(), output, err, errlen);
        if (rc) { remove(output); return rc; }
    }
    t_mux = lap(tl);
    if (trace)
        fprintf(stderr, "[vcpenc] trace open=%.3f pinned_alloc=%.3f read_decode=%.3f session=%.3f gpu=%.3f mux_write=%.3f s\n",
                t_open, t_alloc, t_read, t_create, t_gpu, t_mux);
    const double sec = std::chrono::duration<double>(clock::now() - t0).count();
    fprintf(stderr, "[vcpenc] frames=%ld size=%dx%d fps=%.1f elapsed=%.3fs output=%s\n", total, p.width, p.height,
            sec > 0 ? total / sec : 0.0, sec, output);
    return VCPENC_OK;
}
