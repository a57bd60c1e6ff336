// ISO BMFF (MP4) writer for one AVC or HEVC video track: ftyp / moov / mdat with
// stsd(avc1+avcC | hvc1+hvcC), stts, stss, stsc, stsz, stco|co64; `moov` first when faststart.
//
// Replaces libavformat's `mov` muxer (and its second pass for `-movflags +faststart`) inside
// the ffmpeg child the reference spawns (/root/reference/cmd/consumer.go:376-382; output is
// always *.mp4, /root/reference/cmd/producer.go:417-425).
#include <cstdio>
#include <cstring>

#include "host_bits.h"
#include "host_util.h"

namespace vcp {

std::vector<NalRef> split_annexb(const uint8_t* d, size_t n) {
    std::vector<NalRef> out;
    size_t i = 0, start = (size_t)-1;
    while (i + 3 <= n) {
        if (d[i] == 0 && d[i + 1] == 0 && d[i + 2] == 1) {
            if (start != (size_t)-1) {
                size_t e = i;
                while (e > start && d[e - 1] == 0) e--;
                out.push_back({d + start, e - start});
            }
            start = i + 3;
            i += 3;
        } else i++;
    }
    if (start != (size_t)-1 && start <= n) out.push_back({d + start, n - start});
    return out;
}

namespace {

struct Box {
    std::vector<uint8_t> d;
    void u8(uint32_t v) { d.push_back((uint8_t)v); }
    void u16(uint32_t v) { u8(v >> 8); u8(v); }
    void u24(uint32_t v) { u8(v >> 16); u8(v >> 8); u8(v); }
    void u32(uint32_t v) { u8(v >> 24); u8(v >> 16); u8(v >> 8); u8(v); }
    void u64(uint64_t v) { u32((uint32_t)(v >> 32)); u32((uint32_t)v); }
    void tag(const char* t) { d.insert(d.end(), t, t + 4); }
    void bytes(const void* p, size_t n) { const uint8_t* q = (const uint8_t*)p; d.insert(d.end(), q, q + n); }
    void zeros(size_t n) { d.insert(d.end(), n, 0); }
    size_t begin(const char* t) { size_t at = d.size(); u32(0); tag(t); return at; }
    void end(size_t at) {
        uint32_t sz = (uint32_t)(d.size() - at);
        d[at] = sz >> 24; d[at + 1] = sz >> 16; d[at + 2] = sz >> 8; d[at + 3] = sz;
    }
    void full(uint32_t version, uint32_t flags) { u8(version); u24(flags); }
};

void unity_matrix(Box& b) {
    const uint32_t m[9] = {0x00010000, 0, 0, 0, 0x00010000, 0, 0, 0, 0x40000000};
    for (uint32_t v : m) b.u32(v);
}

// HEVCDecoderConfigurationRecord (ISO/IEC 14496-15 8.3.3.1): profile / tier / level copied out of the SPS,
// then one complete array per parameter-set type
void hvcc_box(Box& b, const ParamSets& ps) {
    std::vector<uint8_t> sps;            // SPS without emulation prevention bytes
    for (size_t i = 0, z = 0; i < ps.sps.size(); i++) {
        const uint8_t v = ps.sps[i];
        if (z >= 2 && v == 3) { z = 0; continue; }
        sps.push_back(v);
        z = v == 0 ? z + 1 : 0;
    }
    size_t c = b.begin("hvcC");
    b.u8(1);
    // sps: 2 bytes NAL header, 1 byte vps id / sub layers / nesting, then general profile_tier_level: 12 bytes
    for (int i = 0; i < 12; i++) b.u8(sps.size() > (size_t)(3 + i) ? sps[3 + i] : 0);
    b.u16(0xF000);                       // min_spatial_segmentation_idc 0
    b.u8(0xFC);                          // parallelismType 0 (unknown)
    b.u8(0xFC | 1);                      // chroma_format_idc 4:2:0
    b.u8(0xF8); b.u8(0xF8);              // bit depth luma / chroma minus 8
    b.u16(0);                            // avgFrameRate: unspecified
    b.u8((0 << 6) | (1 << 3) | (1 << 2) | 3);   // constantFrameRate 0, numTemporalLayers 1, temporalIdNested 1, 4-byte lengths
    b.u8(3);
    const std::vector<uint8_t>* arr[3] = {&ps.vps, &ps.sps, &ps.pps};
    const int types[3] = {32, 33, 34};
    for (int i = 0; i < 3; i++) {
        b.u8(0x80 | types[i]);           // array_completeness 1
        b.u16(1);
        b.u16((uint32_t)arr[i]->size()); b.bytes(arr[i]->data(), arr[i]->size());
    }
    b.end(c);
}

std::vector<uint8_t> build_moov(const vcpenc_params& p, const ParamSets& ps,
                                const std::vector<Mp4Sample>& samples, uint64_t chunk_offset) {
    const std::vector<uint8_t>&sps = ps.sps, &pps = ps.pps;
    const bool hevc = p.codec == VCPENC_CODEC_HEVC;
    const uint32_t n = (uint32_t)samples.size();
    const uint32_t mts = (uint32_t)p.fps_num, delta = (uint32_t)p.fps_den;   // media timescale / sample delta
    const uint64_t mdur = (uint64_t)n * delta;
    const uint32_t movie_ts = 1000;
    const uint64_t dur = mts ? mdur * movie_ts / mts : 0;
    Box b;
    size_t moov = b.begin("moov");
    {
        size_t a = b.begin("mvhd"); b.full(0, 0);
        b.u32(0); b.u32(0); b.u32(movie_ts); b.u32((uint32_t)dur);
        b.u32(0x00010000); b.u16(0x0100); b.u16(0); b.u32(0); b.u32(0);
        unity_matrix(b);
        b.zeros(24); b.u32(2);
        b.end(a);
    }
    size_t trak = b.begin("trak");
    {
        size_t a = b.begin("tkhd"); b.full(0, 3);
        b.u32(0); b.u32(0); b.u32(1); b.u32(0); b.u32((uint32_t)dur);
        b.u32(0); b.u32(0); b.u16(0); b.u16(0); b.u16(0); b.u16(0);
        unity_matrix(b);
        b.u32((uint32_t)p.width << 16); b.u32((uint32_t)p.height << 16);
        b.end(a);
    }
    size_t mdia = b.begin("mdia");
    {
        size_t a = b.begin("mdhd"); b.full(0, 0);
        b.u32(0); b.u32(0); b.u32(mts); b.u32((uint32_t)mdur); b.u16(0x55C4); b.u16(0);
        b.end(a);
        a = b.begin("hdlr"); b.full(0, 0);
        b.u32(0); b.tag("vide"); b.zeros(12); b.bytes("VideoHandler", 13);
        b.end(a);
    }
    size_t minf = b.begin("minf");
    {
        size_t a = b.begin("vmhd"); b.full(0, 1); b.zeros(8); b.end(a);
        a = b.begin("dinf");
        size_t d2 = b.begin("dref"); b.full(0, 0); b.u32(1);
        size_t u = b.begin("url "); b.full(0, 1); b.end(u);
        b.end(d2); b.end(a);
    }
    size_t stbl = b.begin("stbl");
    {
        size_t a = b.begin("stsd"); b.full(0, 0); b.u32(1);
        size_t e = b.begin(hevc ? "hvc1" : "avc1");
        b.zeros(6); b.u16(1);
        b.zeros(16);
        b.u16((uint32_t)p.width); b.u16((uint32_t)p.height);
        b.u32(0x00480000); b.u32(0x00480000); b.u32(0); b.u16(1);
        b.zeros(32);
        b.u16(0x0018); b.u16(0xFFFF);
        if (hevc) hvcc_box(b, ps);
        else {
            size_t c = b.begin("avcC");
            b.u8(1); b.u8(sps.size() > 1 ? sps[1] : 66); b.u8(sps.size() > 2 ? sps[2] : 0); b.u8(sps.size() > 3 ? sps[3] : 40);
            b.u8(0xFF); b.u8(0xE1); b.u16((uint32_t)sps.size()); b.bytes(sps.data(), sps.size());
            b.u8(1); b.u16((uint32_t)pps.size()); b.bytes(pps.data(), pps.size());
            b.end(c);
        }
        b.end(e); b.end(a);

        a = b.begin("stts"); b.full(0, 0); b.u32(1); b.u32(n); b.u32(delta); b.end(a);
        a = b.begin("stss"); b.full(0, 0);
        uint32_t nsync = 0;
        for (const auto& s : samples) nsync += s.sync;
        b.u32(nsync);
        for (uint32_t i = 0; i < n; i++) if (samples[i].sync) b.u32(i + 1);
        b.end(a);
        a = b.begin("stsc"); b.full(0, 0); b.u32(1); b.u32(1); b.u32(n); b.u32(1); b.end(a);
        a = b.begin("stsz"); b.full(0, 0); b.u32(0); b.u32(n);
        for (const auto& s : samples) b.u32(s.size);
        b.end(a);
        if (chunk_offset > 0xFFFFFFFFull) { a = b.begin("co64"); b.full(0, 0); b.u32(1); b.u64(chunk_offset); b.end(a); }
        else { a = b.begin("stco"); b.full(0, 0); b.u32(1); b.u32((uint32_t)chunk_offset); b.end(a); }
    }
    b.end(stbl); b.end(minf); b.end(mdia); b.end(trak); b.end(moov);
    return b.d;
}

}  // namespace

int write_mp4(const vcpenc_params& p, const ParamSets& ps,
              const std::vector<Mp4Sample>& samples, const uint8_t* mdat, uint64_t mdat_len, const char* path,
              char* err, size_t errlen) {
    Box ftyp;
    size_t a = ftyp.begin("ftyp");
    ftyp.tag("isom"); ftyp.u32(0x200); ftyp.tag("isom"); ftyp.tag("iso2"); ftyp.tag(p.codec == VCPENC_CODEC_HEVC ? "hvc1" : "avc1"); ftyp.tag("mp41");
    ftyp.end(a);
    const bool big = mdat_len + 8 > 0xFFFFFFFFull;
    const uint64_t mdat_hdr = big ? 16 : 8;
    uint64_t chunk_off;
    std::vector<uint8_t> moov;
    if (p.faststart) {
        // moov size depends on stco vs co64; iterate once
        moov = build_moov(p, ps, samples, 0);
        chunk_off = ftyp.d.size() + moov.size() + mdat_hdr;
        std::vector<uint8_t> m2 = build_moov(p, ps, samples, chunk_off);
        if (m2.size() != moov.size()) { chunk_off = ftyp.d.size() + m2.size() + mdat_hdr; m2 = build_moov(p, ps, samples, chunk_off); }
        moov.swap(m2);
    } else {
        chunk_off = ftyp.d.size() + mdat_hdr;
        moov = build_moov(p, ps, samples, chunk_off);
    }
    FILE* f = fopen(path, "wb");
    if (!f) { set_err(err, errlen, "cannot create %s", path); return VCPENC_E_IO; }
    bool ok = fwrite(ftyp.d.data(), 1, ftyp.d.size(), f) == ftyp.d.size();
    Box mh;
    if (big) { mh.u32(1); mh.tag("mdat"); mh.u64(mdat_len + 16); } else { mh.u32((uint32_t)(mdat_len + 8)); mh.tag("mdat"); }
    if (p.faststart) ok = ok && fwrite(moov.data(), 1, moov.size(), f) == moov.size();
    ok = ok && fwrite(mh.d.data(), 1, mh.d.size(), f) == mh.d.size();
    ok = ok && (mdat_len == 0 || fwrite(mdat, 1, mdat_len, f) == mdat_len);
    if (!p.faststart) ok = ok && fwrite(moov.data(), 1, moov.size(), f) == moov.size();
    ok = (fclose(f) == 0) && ok;
    if (!ok) { set_err(err, errlen, "short write to %s", path); return VCPENC_E_IO; }
    return VCPENC_OK;
}

}  // namespace vcp

using namespace vcp;

// Annex-B (with per-frame index) -> MP4.  Parameter sets go to avcC / hvcC and are dropped from the
// samples; every other NAL gets a 4-byte length prefix.
extern "C" int vcpenc_mux_mp4(const vcpenc_params* p, const uint8_t* annexb, size_t len, const vcpenc_frame_info* info,
                              int nframes, const char* path, char* err, size_t errlen) {
    if (!p || !annexb || !info || nframes < 1 || !path) { set_err(err, errlen, "bad arguments"); return VCPENC_E_ARGS; }
    std::vector<uint8_t> mdat;
    ParamSets ps;
    std::vector<Mp4Sample> samples;
    mdat.reserve(len + (size_t)nframes * 8);
    for (int i = 0; i < nframes; i++) {
        if (info[i].offset + info[i].size > len) { set_err(err, errlen, "frame index out of range"); return VCPENC_E_ARGS; }
        const auto nals = split_annexb(annexb + info[i].offset, info[i].size);
        Mp4Sample s{mdat.size(), 0, info[i].is_idr != 0};
        for (const auto& nal : nals) {
            if (!nal.n) continue;
            if (ps.take(p->codec, nal)) continue;
            const uint32_t n = (uint32_t)nal.n;
            const uint8_t h[4] = {(uint8_t)(n >> 24), (uint8_t)(n >> 16), (uint8_t)(n >> 8), (uint8_t)n};
            mdat.insert(mdat.end(), h, h + 4);
            mdat.insert(mdat.end(), nal.p, nal.p + nal.n);
        }
        s.size = (uint32_t)(mdat.size() - s.offset);
        samples.push_back(s);
    }
    const bool hevc = p->codec == VCPENC_CODEC_HEVC;
    if (hevc && ps.vps.empty()) ps.vps = make_hevc_vps_nal(*p);
    if (ps.sps.empty()) ps.sps = hevc ? make_hevc_sps_nal(*p) : make_sps_nal(*p);
    if (ps.pps.empty()) ps.pps = hevc ? make_hevc_pps_nal(*p) : make_pps_nal(*p);
    return write_mp4(*p, ps, samples, mdat.data(), mdat.size(), path, err, errlen);
}
