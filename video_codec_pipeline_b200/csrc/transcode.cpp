// vcpenc_transcode — the drop-in for runFFmpegWithTimeout
// (/root/reference/cmd/consumer.go:370-394): input path, output path, whitespace-split option
// tokens, a timeout and a cancel flag in; a complete output file (MP4, or raw .h264) or an
// error class out.  Semantics kept from the reference:
//   * `-y`: an existing output is overwritten (:376)
//   * nil error <=> the file is complete on return (:386,393)
//   * deadline  -> "编码超时", parent cancel -> "任务被取消" (:387-392)
//   * on failure the caller removes the partial output (:264); we also do not leave one
//   * stdout/stderr are inherited: progress lines go to stderr in key=value form
// Frames are pulled in chunks of whole GOPs by a reader thread into one of two page-locked buffers
// while the device session (K1..K5) encodes the other; the access units of a chunk go straight into
// the output file (mux_mp4.cpp), with the input's audio beside them; no frame ever takes a CPU encode
// path.  Cancellation and the deadline are polled once per GOP while reading and between the stages of
// a chunk.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
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

// Raw and y4m pictures sit at computable offsets: a chunk is filled by several threads with pread(), each taking
// a share of the pictures (one thread copies ~4 GB/s out of the page cache; the encoder takes 40+ GB/s).
int read_workers() {
    static const int n = [] {
        if (const char* e = getenv("VCPENC_READ_THREADS")) return std::max(1, atoi(e));
        const long c = sysconf(_SC_NPROCESSORS_ONLN);
        return (int)std::max<long>(1, std::min<long>(8, c / 2));
    }();
    return n;
}
// pictures [first, first + count) of `hdr`-prefixed records of `fb` bytes starting at `base`; checks the y4m FRAME marker
int pread_frames(int fd, uint64_t base, size_t hdr, size_t fb, long first, int count, uint8_t* dst, bool* bad_marker) {
    const int nw = std::min(read_workers(), std::max(1, count / 4));
    std::vector<int> got((size_t)nw, 0);
    auto work = [&](int k) {
        const int a = (int)((long long)count * k / nw), b = (int)((long long)count * (k + 1) / nw);
        int n = 0;
        for (int i = a; i < b; i++) {
            const uint64_t off = base + (uint64_t)(first + i) * (hdr + fb);
            if (hdr) {
                char m[8];
                if (pread(fd, m, hdr, (off_t)off) != (ssize_t)hdr) break;
                if (memcmp(m, "FRAME\n", 6)) { *bad_marker = true; break; }
            }
            size_t done = 0;
            while (done < fb) {
                const ssize_t r = pread(fd, dst + (size_t)i * fb + done, fb - done, (off_t)(off + hdr + done));
                if (r <= 0) break;
                done += (size_t)r;
            }
            if (done < fb) break;
            n++;
        }
        got[(size_t)k] = n;
    };
    if (nw == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int k = 0; k < nw; k++) th.emplace_back(work, k);
        for (auto& t : th) t.join();
    }
    // pictures are usable up to the first share that came up short
    int n = 0;
    for (int k = 0; k < nw; k++) {
        const int a = (int)((long long)count * k / nw), b = (int)((long long)count * (k + 1) / nw);
        n += got[(size_t)k];
        if (got[(size_t)k] < b - a) break;
    }
    return n;
}

struct Y4mSource : FrameSource {
    FILE* f = nullptr;
    uint64_t data0 = 0;     // offset of the first FRAME marker
    long next = 0;          // next picture (positional reads)
    bool plain = true;      // every marker so far was the bare "FRAME\n": offsets are computable
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
        data0 = (uint64_t)ftello(f);
        struct stat sb;
        if (stat(path, &sb) == 0 && sb.st_size > 0) {
            est_frames = (long)((size_t)sb.st_size / (fbytes() + 6)) + 1;
            if (S_ISREG(sb.st_mode) && (uint64_t)sb.st_size > data0) exact_frames = (long)(((uint64_t)sb.st_size - data0) / (fbytes() + 6));   // bare FRAME markers assumed; a short chunk falls back
        }
        return VCPENC_OK;
    }
    int read(uint8_t* dst, int max, char* err, size_t errlen) override {
        const size_t fb = fbytes();
        int n = 0;
        if (plain) {
            bool bad = false;
            n = pread_frames(fileno(f), data0, 6, fb, next, max, dst, &bad);
            next += n;
            if (!bad) return n;
            // a FRAME marker with parameters: continue with the sequential parser from the first picture not read yet
            plain = false;
            if (fseeko(f, (off_t)(data0 + (uint64_t)next * (6 + fb)), SEEK_SET) != 0) { set_err(err, errlen, "y4m: seek failed"); return -VCPENC_E_IO; }
        }
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
    long next = 0;
    ~RawSource() override { if (f) fclose(f); }
    int open(const char* path, const vcpenc_params& p, int pixfmt, char* err, size_t errlen) {
        if (p.in_width <= 0 || p.in_height <= 0) { set_err(err, errlen, "raw input needs -s WxH"); return VCPENC_E_FORMAT; }
        fmt = pixfmt;
        width = p.in_width; height = p.in_height; fps_num = p.fps_num; fps_den = p.fps_den;
        f = fopen(path, "rb");
        if (!f) { set_err(err, errlen, "cannot open %s", path); return VCPENC_E_IO; }
        struct stat sb;
        if (stat(path, &sb) == 0 && sb.st_size > 0) {
            est_frames = (long)((size_t)sb.st_size / fbytes()) + 1;
            if (S_ISREG(sb.st_mode)) exact_frames = (long)((size_t)sb.st_size / fbytes());
        }
        return VCPENC_OK;
    }
    int read(uint8_t* dst, int max, char*, size_t) override {
        bool bad = false;
        const int n = pread_frames(fileno(f), 0, 0, fbytes(), next, max, dst, &bad);
        next += n;
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

}  // namespace


static thread_local int t_device = 0;

// encoder.cu: free / total bytes of a device (0 on failure)
size_t vcp_device_free_bytes(int device, size_t* total);

// ---------------------------------------------------------------------------------------------
// Reuse across tasks.  Measured on the task-flow harness: session create + destroy 0.1-1.4 s and
// cudaMallocHost 0.1-0.3 s per task against 0.13 s of GPU work for a 120-frame 1080p clip.
//   * sessions (all device buffers of one geometry / preset) live in a PROCESS-WIDE pool keyed by
//     (device, parameters): a worker checks one out for a task and hands it back; idle sessions are
//     destroyed least-recently-used first when a create fails for lack of memory, when the pool holds
//     more than kPoolSlots, or when the idle ones exceed half of the device's memory;
//   * the page-locked chunk buffers come from a process-wide pool too (smallest idle buffer that is large enough).
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int kPoolSlots = 8;
struct PoolEntry {
    vcpenc_session* ses = nullptr;
    vcpenc_params key{};
    int max_frames = 0, device = -1;
    size_t bytes = 0;
    unsigned long stamp = 0;
    bool in_use = false;
};
std::mutex g_pool_mu;
std::vector<PoolEntry> g_pool;
unsigned long g_pool_clock = 0;

// page-locked chunk buffers: process-wide as well (a Go caller's goroutines move between OS threads, and a worker
// that comes back on another thread must still find its buffers); idle buffers beyond kPinnedIdleMax are freed
struct PinnedEntry { uint8_t* p = nullptr; size_t bytes = 0; bool in_use = false; unsigned long stamp = 0; };
std::vector<PinnedEntry> g_pinned;
constexpr size_t kPinnedIdleMax = (size_t)16 << 30;
bool cache_enabled() { static const bool on = getenv("VCPENC_NO_CACHE") == nullptr; return on; }
bool same_key(vcpenc_params a, vcpenc_params b) { a.first_gop = b.first_gop = 0; return memcmp(&a, &b, sizeof a) == 0; }

// destroy the least recently used idle session of `device` (any device if < 0); pool lock held
bool evict_one_locked(int device) {
    int victim = -1;
    for (int i = 0; i < (int)g_pool.size(); i++)
        if (!g_pool[i].in_use && (device < 0 || g_pool[i].device == device) && (victim < 0 || g_pool[i].stamp < g_pool[victim].stamp)) victim = i;
    if (victim < 0) return false;
    vcpenc_session_destroy(g_pool[victim].ses);
    g_pool.erase(g_pool.begin() + victim);
    return true;
}
}  // namespace

// session of `want` frames on `device` for parameters p: an idle one from the pool, else new
int acquire_session(const vcpenc_params& p, int device, int want, vcpenc_session** out, char* err, size_t errlen) {
    if (cache_enabled()) {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        for (auto& c : g_pool)
            if (!c.in_use && c.device == device && c.max_frames >= want && same_key(c.key, p)) {
                c.in_use = true; c.stamp = ++g_pool_clock; *out = c.ses;
                return VCPENC_OK;
            }
        // an idle session with the same parameters but too small is of no further use
        for (size_t i = 0; i < g_pool.size(); i++)
            if (!g_pool[i].in_use && g_pool[i].device == device && same_key(g_pool[i].key, p)) { vcpenc_session_destroy(g_pool[i].ses); g_pool.erase(g_pool.begin() + i); break; }
    }
    size_t total = 0;
    const size_t before = vcp_device_free_bytes(device, &total);
    int rc;
    for (;;) {
        rc = vcpenc_session_create(&p, device, want, out, err, errlen);
        if (rc != VCPENC_E_CUDA) break;
        // most likely out of memory: give back what idle sessions hold on this device and try again
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (!evict_one_locked(device)) break;
    }
    if (rc) return rc;
    if (cache_enabled()) {
        const size_t after = vcp_device_free_bytes(device, nullptr);
        std::lock_guard<std::mutex> lk(g_pool_mu);
        PoolEntry e;
        e.ses = *out; e.key = p; e.max_frames = want; e.device = device; e.in_use = true; e.stamp = ++g_pool_clock;
        e.bytes = before > after ? before - after : 0;
        g_pool.push_back(e);
        while ((int)g_pool.size() > kPoolSlots && evict_one_locked(-1)) {}
    }
    return VCPENC_OK;
}

// hand a session back; `ok` false: its state is not trusted after a failure
void release_session(vcpenc_session* ses, bool ok) {
    if (!ses) return;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (size_t i = 0; i < g_pool.size(); i++)
        if (g_pool[i].ses == ses) {
            if (!ok) { vcpenc_session_destroy(ses); g_pool.erase(g_pool.begin() + i); return; }
            g_pool[i].in_use = false; g_pool[i].stamp = ++g_pool_clock;
            // idle sessions may hold at most half of the device
            const int device = g_pool[i].device;
            size_t total = 0;
            vcp_device_free_bytes(device, &total);
            for (;;) {
                size_t idle = 0;
                for (const auto& c : g_pool) if (!c.in_use && c.device == device) idle += c.bytes;
                if (total == 0 || idle <= total / 2 || !evict_one_locked(device)) break;
            }
            return;
        }
    vcpenc_session_destroy(ses);   // not pooled (VCPENC_NO_CACHE)
}

uint8_t* acquire_pinned(size_t bytes) {
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        int best = -1;
        for (int i = 0; i < (int)g_pinned.size(); i++)
            if (!g_pinned[i].in_use && g_pinned[i].bytes >= bytes && (best < 0 || g_pinned[i].bytes < g_pinned[best].bytes)) best = i;
        if (best >= 0) { g_pinned[best].in_use = true; g_pinned[best].stamp = ++g_pool_clock; return g_pinned[best].p; }
    }
    uint8_t* p = (uint8_t*)vcpenc_host_alloc(bytes);
    if (!p) {   // give idle buffers back to the system and try once more
        std::vector<uint8_t*> drop;
        {
            std::lock_guard<std::mutex> lk(g_pool_mu);
            for (size_t i = 0; i < g_pinned.size();) if (!g_pinned[i].in_use) { drop.push_back(g_pinned[i].p); g_pinned.erase(g_pinned.begin() + i); } else i++;
        }
        for (uint8_t* q : drop) vcpenc_host_free(q);
        p = (uint8_t*)vcpenc_host_alloc(bytes);
    }
    if (p && cache_enabled()) {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        PinnedEntry e; e.p = p; e.bytes = bytes; e.in_use = true; e.stamp = ++g_pool_clock;
        g_pinned.push_back(e);
    }
    return p;
}
void release_pinned(uint8_t* p) {
    if (!p) return;
    std::vector<uint8_t*> drop;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        bool pooled = false;
        for (auto& e : g_pinned) if (e.p == p) { e.in_use = false; e.stamp = ++g_pool_clock; pooled = true; }
        if (!pooled) drop.push_back(p);
        for (;;) {   // oldest idle buffers go first
            size_t idle = 0; int victim = -1;
            for (int i = 0; i < (int)g_pinned.size(); i++)
                if (!g_pinned[i].in_use) { idle += g_pinned[i].bytes; if (victim < 0 || g_pinned[i].stamp < g_pinned[victim].stamp) victim = i; }
            if (idle <= kPinnedIdleMax || victim < 0) break;
            drop.push_back(g_pinned[victim].p);
            g_pinned.erase(g_pinned.begin() + victim);
        }
    }
    for (uint8_t* q : drop) vcpenc_host_free(q);
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
    std::vector<uint8_t*> drop;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        while (evict_one_locked(-1)) {}
        for (size_t i = 0; i < g_pinned.size();) if (!g_pinned[i].in_use) { drop.push_back(g_pinned[i].p); g_pinned.erase(g_pinned.begin() + i); } else i++;
    }
    for (uint8_t* q : drop) vcpenc_host_free(q);
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
    // VCPENC_TRACE=1: where the wall time of a task goes (open / pinned alloc / waiting for the reader / session / GPU / mux)
    const bool trace = getenv("VCPENC_TRACE") != nullptr;
    double t_open = 0, t_alloc = 0, t_wait = 0, t_create = 0, t_gpu = 0, t_mux = 0;
    auto lap = [&](clock::time_point& from) { const auto now = clock::now(); const double d = std::chrono::duration<double>(now - from).count(); from = now; return d; };
    auto tl = t0;
    if (!input || !output) { set_err(err, errlen, "bad arguments"); return VCPENC_E_ARGS; }
    vcpenc_params p;
    int rc = vcpenc_parse_args(argc, argv, &p, err, errlen);
    if (rc) return rc;
    if (vcpenc_device_count() <= 0) { set_err(err, errlen, "no CUDA device available (libvcpenc has no CPU fallback)"); return VCPENC_E_NODEVICE; }
    if ((p.maxrate > 0) != (p.bufsize > 0) && !getenv("VCPENC_QUIET")) {
        // like libx264: a VBV needs both numbers; -maxrate alone still caps the GOP budget of -b:v
        static std::atomic<bool> once{false};
        if (!once.exchange(true)) fprintf(stderr, "[vcpenc] warning: VBV needs both -maxrate and -bufsize; the buffer model is off\n");
    }

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
        rc = open_container_source(input, p.drop_audio != 0, p.audio_bitrate, &src, err, errlen);
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
    vcpenc_params pkey = p;            // session key: what the kernels see (not the muxer's / front end's settings)
    pkey.faststart = 0; pkey.audio_bitrate = 0; pkey.drop_audio = 0;

    const size_t sfb = src->fbytes();
    const size_t fb = std::max(fbytes(p.width, p.height), sfb);
    // chunk: whole GOPs, at most ~3 GiB of raw frames per pass and buffer (17 closed GOPs of 1080p).  Measured on a 1 920
    // picture y4m with two worker threads (profiles/r02_notes.md): 1.5 GiB 6 584 fps, 2 GiB 7 606, 3 GiB 7 691, 6 GiB 6 346 --
    // larger chunks fill the GOP-group chains better but leave the reader nothing to overlap with.
    const int ndev = task_gpus();
    size_t chunk_bytes = ((size_t)3 << 30) * (size_t)ndev;
    if (const char* e = getenv("VCPENC_CHUNK_BYTES")) { const long long v = atoll(e); if (v > 0) chunk_bytes = (size_t)v; }   // tests: force many chunks
    int chunk = (int)std::max<size_t>(1, chunk_bytes / fb);
    // short clips: do not page-lock more host memory than the input can fill.  The picture count is exact for raw /
    // y4m files and an estimate (duration x frame rate) for containers; unknown (0) keeps the full chunk.
    const long est = src->est_frames;
    if (est > 0) chunk = (int)std::min<long>((long)chunk, est);
    chunk = std::max(p.gop, (chunk + p.gop - 1) / p.gop * p.gop);
    const int nbuf = (est > 0 && est <= chunk) ? 1 : 2;   // one chunk holds everything: nothing to overlap
    uint8_t* fbuf[2] = {nullptr, nullptr};
    for (int i = 0; i < nbuf; i++) {
        const size_t need = (size_t)chunk * sfb;
        fbuf[i] = acquire_pinned(need);
        if (!fbuf[i]) {
            for (int k = 0; k < i; k++) release_pinned(fbuf[k]);
            set_err(err, errlen, "cannot allocate %zu bytes of pinned host memory", need);
            return VCPENC_E_CUDA;
        }
    }
    t_alloc = lap(tl);

    struct Shard {
        int device = 0; vcpenc_session* ses = nullptr;
        // bitstream of the shard's pictures: NOT a std::vector -- resize() would zero-fill hundreds of megabytes per task
        // (measured: 0.17 s of a 0.30 s 300-frame task); grown on VCPENC_E_OVERFLOW
        std::unique_ptr<uint8_t[]> bits; size_t bits_cap = 0;
        void need_bits(size_t n) { if (bits_cap < n) { bits.reset(new uint8_t[n]); bits_cap = n; } }
        std::vector<vcpenc_frame_info> info;
        int f0 = 0, n = 0, g0 = 0; size_t len = 0; int rc = 0; char err[256] = {0};
    };
    std::vector<Shard> shards((size_t)ndev);
    for (int d = 0; d < ndev; d++) shards[d].device = (t_device + d) % std::max(1, vcpenc_device_count());
    const bool raw_out = ends_with(outp, ".h264") || ends_with(outp, ".264") || ends_with(outp, ".h265") || ends_with(outp, ".265") || ends_with(outp, ".hevc");
    Mp4Writer mp4;
    FILE* rawf = nullptr;
    bool writer_open = false;

    // ---- reader thread: fills the chunk buffers one GOP at a time, polling the stop conditions in between ----
    std::mutex mu;
    std::condition_variable cv;
    int filled[2] = {-1, -1};          // frames in buffer i; -1: free for the reader
    bool last[2] = {false, false};     // buffer i holds the end of the input
    // upload-while-reading (inputs whose picture count is known): pictures of buffer i read so far / the reader is done with it
    volatile long progress[2] = {0, 0};
    volatile int finished[2] = {0, 0};
    int read_rc = 0; char read_err[256] = {0};
    std::atomic<int> stop{0};          // VCPENC_E_CANCELLED / VCPENC_E_TIMEOUT once a stop condition is seen
    bool quit = false;
    auto stop_reason = [&]() -> int {
        if (stop.load()) return stop.load();
        if (cancel && *cancel) { stop.store(VCPENC_E_CANCELLED); return VCPENC_E_CANCELLED; }
        if (timeout_ms > 0 && std::chrono::duration_cast<std::chrono::milliseconds>(clock::now() - t0).count() > timeout_ms) { stop.store(VCPENC_E_TIMEOUT); return VCPENC_E_TIMEOUT; }
        return 0;
    };
    std::thread reader([&] {
        for (int i = 0;; i = (i + 1) % nbuf) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return filled[i] < 0 || quit; });
                if (quit) return;
            }
            int n = 0, rcr = 0;
            bool eof = false;
            while (n < chunk && !eof) {
                if (stop_reason()) break;
                const int want = std::min(p.gop, chunk - n);
                const int k = src->read(fbuf[i] + (size_t)n * sfb, want, read_err, sizeof read_err);
                if (k < 0) { rcr = -k; break; }
                n += k;
                std::atomic_thread_fence(std::memory_order_release);
                progress[i] = n;
                if (k < want) eof = true;
            }
            std::atomic_thread_fence(std::memory_order_release);
            finished[i] = 1;
            if (eof && !rcr) { const int ra = src->finish_audio(read_err, sizeof read_err); if (ra) rcr = ra; }
            std::lock_guard<std::mutex> lk(mu);
            filled[i] = n; last[i] = eof || rcr || stop.load(); read_rc = rcr;
            cv.notify_all();
            if (last[i]) return;
        }
    });
    auto fail = [&](int code) {
        { std::lock_guard<std::mutex> lk(mu); quit = true; stop.store(stop.load() ? stop.load() : code); cv.notify_all(); }
        if (reader.joinable()) reader.join();
        for (auto& sh : shards) if (sh.ses) { release_session(sh.ses, false); sh.ses = nullptr; }
        if (rawf) { fclose(rawf); rawf = nullptr; }
        if (writer_open && !raw_out) mp4.abandon();
        remove(output);
        for (int i = 0; i < nbuf; i++) release_pinned(fbuf[i]);
        return code;
    };
    auto stop_message = [&](int code) {
        if (code == VCPENC_E_CANCELLED) set_err(err, errlen, "任务被取消");
        else set_err(err, errlen, "编码超时 (>%dms)", timeout_ms);
    };
    // audio access units the front end has produced so far go into the file behind the video of the chunk
    auto drain_audio = [&]() -> int {
        if (!src->audio.present || raw_out) return 0;
        std::vector<uint8_t> data; std::vector<uint32_t> sizes;
        { std::lock_guard<std::mutex> lk(src->audio_mu); data.swap(src->audio.data); sizes.swap(src->audio.sizes); }
        size_t o = 0;
        for (uint32_t sz : sizes) { if (mp4.audio_frame(data.data() + o, sz)) return VCPENC_E_IO; o += sz; }
        return 0;
    };

    long total = 0;
    int gop_index = 0;
    // shards of a chunk of n pictures: closed GOPs, contiguous ranges per device; sessions on first use
    auto prepare = [&](int n) -> int {
        const int ngops = (n + p.gop - 1) / p.gop;
        const int used = std::min(ndev, ngops);
        for (int d = 0; d < ndev; d++) {
            Shard& sh = shards[d];
            const int ga = d < used ? (int)((long long)ngops * d / used) : 0, gb = d < used ? (int)((long long)ngops * (d + 1) / used) : 0;
            sh.f0 = ga * p.gop; sh.n = std::min(n, gb * p.gop) - sh.f0; sh.g0 = gop_index + ga; sh.rc = 0; sh.len = 0;
            if (sh.n <= 0) { sh.n = 0; continue; }
            if (!sh.ses) {
                pkey.first_gop = 0;
                const int want = std::min((chunk / p.gop + used - 1) / used * p.gop, std::max(sh.n, 1));
                const int rcs = acquire_session(pkey, sh.device, std::max(want, sh.n), &sh.ses, err, errlen);
                if (rcs) return rcs;
            }
            sh.need_bits((size_t)sh.n * fb / 4 + (1 << 20));
            if (sh.info.size() < (size_t)sh.n) sh.info.resize((size_t)sh.n);
        }
        return 0;
    };
    for (int bi = 0;; bi = (bi + 1) % nbuf) {
        int n = 0;
        bool is_last = false;
        // Upload and encode while reading: with the picture count known (raw / y4m files) and one device, the copies of this
        // chunk are queued GOP by GOP as the reader delivers them and the encode of every GOP group right behind its
        // copies, so the H2D transfer hides inside the file read and the first groups are done when the last one arrives.
        // A chunk that comes up short is uploaded and encoded again as it is.
        bool early = false;
        if (ndev == 1 && src->exact_frames > total && !getenv("VCPENC_NO_EARLY_UPLOAD")) {
            const int n_exp = (int)std::min<long>((long)chunk, src->exact_frames - total);
            rc = prepare(n_exp);
            if (rc) return fail(rc);
            t_create += lap(tl);
            Shard& sh = shards[0];
            vcpenc_session_set_first_gop(sh.ses, sh.g0);
            const int rg = vcpenc_session_encode_gated(sh.ses, fbuf[bi], n_exp, &progress[bi], &finished[bi], sh.err, sizeof sh.err);
            if (rg == VCPENC_OK) early = true;
            else if (rg != VCPENC_E_CANCELLED) { set_err(err, errlen, "%s", sh.err); return fail(rg); }
        }
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return filled[bi] >= 0; });
            n = filled[bi]; is_last = last[bi];
        }
        t_wait += lap(tl);
        if (int sr = stop_reason()) { stop_message(sr); return fail(sr); }
        if (read_rc) { set_err(err, errlen, "%s", read_err); return fail(read_rc); }
        if (n > 0) {
            if (early && n != shards[0].n) early = false;   // the chunk came up short (or long): upload it again as it is
            const int ngops = (n + p.gop - 1) / p.gop;
            const int used = std::min(ndev, ngops);
            if (!early) { rc = prepare(n); if (rc) return fail(rc); }
            t_create += lap(tl);
            auto run = [&](Shard& sh) {
                if (!sh.n) return;
                vcpenc_session_set_first_gop(sh.ses, sh.g0);          // idr_pic_id parity continues across ranges and chunks
                // streamed: the pinned chunk buffer outlives the encode, whose GOP groups start as their frames land
                if (!early) {
                    sh.rc = vcpenc_session_upload_async(sh.ses, fbuf[bi] + (size_t)sh.f0 * sfb, sh.n, sh.err, sizeof sh.err);
                    if (!sh.rc) sh.rc = vcpenc_session_encode(sh.ses, nullptr, sh.err, sizeof sh.err);
                }
                if (!sh.rc) {
                    sh.rc = vcpenc_session_download(sh.ses, sh.bits.get(), sh.bits_cap, &sh.len, sh.info.data(), nullptr, sh.err, sizeof sh.err);
                    if (sh.rc == VCPENC_E_OVERFLOW) {
                        sh.need_bits((size_t)sh.n * fb + (1 << 20));
                        sh.rc = vcpenc_session_download(sh.ses, sh.bits.get(), sh.bits_cap, &sh.len, sh.info.data(), nullptr, sh.err, sizeof sh.err);
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
        }
        // the buffer is free again: the reader fills it while this chunk is written out
        { std::lock_guard<std::mutex> lk(mu); filled[bi] = -1; progress[bi] = 0; finished[bi] = 0; cv.notify_all(); }
        if (int sr = stop_reason()) { stop_message(sr); return fail(sr); }
        if (n > 0) {
            if (!writer_open) {
                if (raw_out) {
                    rawf = fopen(output, "wb");
                    if (!rawf) { set_err(err, errlen, "cannot create %s", output); return fail(VCPENC_E_IO); }
                } else {
                    // expected samples (moov reserve of a faststart file): the source's estimate, else what this first chunk suggests
                    const uint64_t ev = est > 0 ? (uint64_t)est + 2 : (is_last ? (uint64_t)n : 0);
                    uint64_t ea = 0;
                    if (src->audio.present && ev && p.fps_num > 0)
                        ea = ev * (uint64_t)p.fps_den * (uint64_t)src->audio.sample_rate / ((uint64_t)p.fps_num * (uint64_t)src->audio.frame_samples) + 64;
                    else if (src->audio.present) ea = 1;
                    rc = mp4.open(output, p, ev, ea, err, errlen);
                    if (rc) return fail(rc);
                }
                writer_open = true;
            }
            for (auto& sh : shards) {            // host concatenation in GOP order
                if (!sh.n) continue;
                if (raw_out) {
                    if (fwrite(sh.bits.get(), 1, sh.len, rawf) != sh.len) { set_err(err, errlen, "short write to %s", output); return fail(VCPENC_E_IO); }
                    continue;
                }
                for (int i = 0; i < sh.n; i++) {
                    if (mp4.video_access_unit(sh.bits.get() + sh.info[i].offset, sh.info[i].size, sh.info[i].is_idr != 0)) { set_err(err, errlen, "short write to %s", output); return fail(VCPENC_E_IO); }
                    if ((i + 1) % p.gop == 0) mp4.end_chunk();
                }
                mp4.end_chunk();
            }
            if (!is_last && drain_audio()) { set_err(err, errlen, "short write to %s", output); return fail(VCPENC_E_IO); }
            total += n;
            gop_index += (n + p.gop - 1) / p.gop;
            t_mux += lap(tl);
        }
        if (is_last) break;
    }
    if (reader.joinable()) reader.join();
    for (auto& sh : shards) { if (sh.ses) release_session(sh.ses, true); sh.ses = nullptr; }
    for (int i = 0; i < nbuf; i++) release_pinned(fbuf[i]);
    if (total == 0) {
        if (rawf) fclose(rawf);
        if (writer_open && !raw_out) mp4.abandon();
        set_err(err, errlen, "input has no frames"); remove(output); return VCPENC_E_FORMAT;
    }
    if (raw_out) {
        if (fclose(rawf) != 0) { set_err(err, errlen, "short write to %s", output); remove(output); return VCPENC_E_IO; }
    } else {
        if (drain_audio()) { set_err(err, errlen, "short write to %s", output); mp4.abandon(); return VCPENC_E_IO; }
        rc = mp4.finish(src->audio.present ? &src->audio : nullptr, err, errlen);
        if (rc) { mp4.abandon(); return rc; }
    }
    t_mux += lap(tl);
    const double sec = std::chrono::duration<double>(clock::now() - t0).count();
    if (trace)
        fprintf(stderr, "[vcpenc] trace open=%.3f pinned_alloc=%.3f wait_for_reader=%.3f session=%.3f gpu=%.3f mux_write=%.3f s\n",
                t_open, t_alloc, t_wait, t_create, t_gpu, t_mux);
    // the reference runs its child with `-loglevel warning` (cmd/consumer.go:376): a successful task prints nothing
    if (trace || getenv("VCPENC_VERBOSE"))
        fprintf(stderr, "[vcpenc] frames=%ld size=%dx%d fps=%.1f elapsed=%.3fs audio=%s output=%s\n", total, p.width, p.height,
                sec > 0 ? total / sec : 0.0, sec, src->audio.present ? (src->audio.copied ? "aac(copy)" : "aac") : "none", output);
    return VCPENC_OK;
}
