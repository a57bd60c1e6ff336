// vcpenc_verify — the `--verify` acceptance check of the reference, without ffprobe.
//
// Reference: verifyOutputFile (/root/reference/cmd/consumer.go:396-419): the file must exist,
// be non-empty, and `ffprobe -select_streams v:0 -show_entries stream=codec_type` must print
// "video".  Here: the container must parse as ISO BMFF with a `moov` holding a track whose
// handler is 'vide' with a non-empty sample table, or be an Annex-B H.264 / HEVC elementary stream
// starting with a parameter set.  Error strings follow the reference's.
#include <sys/stat.h>

#include <cstdio>
#include <cstring>

#include "host_util.h"

using namespace vcp;

namespace {

uint32_t rd32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
uint64_t rd64(const uint8_t* p) { return ((uint64_t)rd32(p) << 32) | rd32(p + 4); }

struct TrackInfo { bool is_video = false; uint32_t samples = 0; bool has_stsd = false; };

// walk boxes in [p, p+n); returns false on malformed sizes
bool walk(const uint8_t* p, size_t n, int depth, TrackInfo* trk, bool* found_video) {
    size_t o = 0;
    while (o + 8 <= n) {
        uint64_t sz = rd32(p + o);
        const uint8_t* t = p + o + 4;
        size_t hdr = 8;
        if (sz == 1) { if (o + 16 > n) return false; sz = rd64(p + o + 8); hdr = 16; }
        else if (sz == 0) sz = n - o;
        if (sz < hdr || o + sz > n) return false;
        const uint8_t* body = p + o + hdr;
        const size_t blen = (size_t)sz - hdr;
        if (!memcmp(t, "trak", 4)) {
            TrackInfo ti;
            if (!walk(body, blen, depth + 1, &ti, found_video)) return false;
            if (ti.is_video && ti.has_stsd && ti.samples > 0) *found_video = true;
        } else if (!memcmp(t, "moov", 4) || !memcmp(t, "mdia", 4) || !memcmp(t, "minf", 4) || !memcmp(t, "stbl", 4)) {
            if (!walk(body, blen, depth + 1, trk, found_video)) return false;
        } else if (trk && !memcmp(t, "hdlr", 4) && blen >= 12) {
            if (!memcmp(body + 8, "vide", 4)) trk->is_video = true;
        } else if (trk && !memcmp(t, "stsd", 4) && blen >= 16) {
            if (rd32(body + 4) >= 1) trk->has_stsd = true;
        } else if (trk && !memcmp(t, "stsz", 4) && blen >= 12) {
            trk->samples = rd32(body + 8);
        }
        o += (size_t)sz;
    }
    return true;
}

}  // namespace

extern "C" int vcpenc_verify(const char* path, char* err, size_t errlen) {
    if (!path) { set_err(err, errlen, "bad arguments"); return VCPENC_E_ARGS; }
    struct stat st;
    if (stat(path, &st) != 0) { set_err(err, errlen, "输出文件不存在: %s", path); return VCPENC_E_IO; }
    if (st.st_size == 0) { set_err(err, errlen, "输出文件为空"); return VCPENC_E_VERIFY; }
    FILE* f = fopen(path, "rb");
    if (!f) { set_err(err, errlen, "输出文件不存在: %s", path); return VCPENC_E_IO; }
    uint8_t head[16] = {0};
    size_t got = fread(head, 1, sizeof head, f);
    bool video = false;
    if (got >= 8 && !memcmp(head + 4, "ftyp", 4)) {
        // top-level scan: read only box headers, load `moov` fully
        uint64_t o = 0;
        const uint64_t fsz = (uint64_t)st.st_size;
        while (o + 8 <= fsz) {
            uint8_t h[16];
            if (fseeko(f, (off_t)o, SEEK_SET) != 0 || fread(h, 1, 8, f) != 8) break;
            uint64_t sz = rd32(h);
            size_t hdr = 8;
            if (sz == 1) { if (fread(h + 8, 1, 8, f) != 8) break; sz = rd64(h + 8); hdr = 16; }
            else if (sz == 0) sz = fsz - o;
            if (sz < hdr || o + sz > fsz) break;
            if (!memcmp(h + 4, "moov", 4)) {
                std::vector<uint8_t> moov((size_t)(sz - hdr));
                if (fread(moov.data(), 1, moov.size(), f) == moov.size()) walk(moov.data(), moov.size(), 1, nullptr, &video);
                break;
            }
            o += sz;
        }
    } else if (got >= 5 && head[0] == 0 && head[1] == 0 && (head[2] == 1 || (head[2] == 0 && head[3] == 1))) {
        const uint8_t* h = head[2] == 1 ? head + 3 : head + 4;
        const int t = h[0] & 31;                 // H.264 nal_unit_type
        const int t5 = (h[0] >> 1) & 63;         // HEVC nal_unit_type (two-byte header: layer 0, temporal id + 1 != 0)
        video = t == 7 || t == 9 || t == 5 || t == 1 ||
                (!(h[0] & 0x81) && (h[1] & 7) && !(h[1] >> 3) && (t5 == 32 || t5 == 33 || t5 == 35 || (t5 >= 16 && t5 <= 21) || t5 <= 9));
    }
    fclose(f);
    if (!video) { set_err(err, errlen, "无有效视频流"); return VCPENC_E_VERIFY; }
    return VCPENC_OK;
}
